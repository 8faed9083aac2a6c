// exp_store.cu -- round-2 GPU experiment (not product code, not a bench number): what bounds compute_inputs once
// its gathers are cheap?  tools/exp_gather2 shows every K1-like variant at ~3.0 ms whatever the gather costs:
// 15 GB of planar stores (16 channels x 7 timesteps per pair) at ~5 TB/s, while a plain fill runs at 7.5 TB/s.
// This program times the STORE PATTERN of that kernel: 16 pairs 1088x1920, 7 timesteps x 16 fp32 planes, for
//   tile shapes   32x8 / 64x4 / 128x2 pixels per 256-thread CTA (contiguous bytes per plane row: 128 .. 512)
//   pixels/thread 1 (STG.32) / 2 (STG.64) / 4 (STG.128): a warp then writes 128 / 256 / 512 contiguous bytes
//   store flavour st.global.cs (streaming, shipped) / st.global (default) / st.global.wt
//   NHWC          one pixel's 16 channels contiguous (2 x STG.256 per thread and timestep)
// with and without the real kernel's gathers (uint8 2x2 entries, one LDG.128 per sample) in front of the stores,
// plus a plain grid-stride fill of the same bytes as the ceiling.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/exp_store tools/exp_store.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int H = 1088, W = 1920, B = 16, N = 7, C = 16;
constexpr long long NPX = (long long)H * W;

__device__ __forceinline__ unsigned hash_u(unsigned a) {
    a ^= a >> 16; a *= 0x7feb352dU; a ^= a >> 15; a *= 0x846ca68bU; a ^= a >> 16; return a;
}
__device__ __forceinline__ float randn_(unsigned k) {
    float u1 = (hash_u(k * 2 + 1) >> 8) * (1.0f / 16777216.0f) + 1e-7f;
    float u2 = (hash_u(k * 2 + 2) >> 8) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
__global__ void make_flow(float* flow, int G, float px, unsigned seed, long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    int x = i % W, y = (i / W) % H; int bc = i / NPX;
    int gw = W / G, gh = H / G;
    float sx = (x + 0.5f) / G - 0.5f, sy = (y + 0.5f) / G - 0.5f;
    sx = fminf(fmaxf(sx, 0.f), gw - 1.f); sy = fminf(fmaxf(sy, 0.f), gh - 1.f);
    int x0 = (int)sx, y0 = (int)sy; int x1 = min(x0 + 1, gw - 1), y1 = min(y0 + 1, gh - 1);
    float fx = sx - x0, fy = sy - y0;
    unsigned base = seed + bc * 1000003u;
    float a = randn_(base + y0 * gw + x0), b = randn_(base + y0 * gw + x1), c = randn_(base + y1 * gw + x0), d = randn_(base + y1 * gw + x1);
    flow[i] = px * ((a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy);
}
__global__ void make_quads(uint4* uq, long long total) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < total) uq[i] = make_uint4(hash_u((unsigned)i), hash_u((unsigned)i + 77u), hash_u((unsigned)i + 991u), 0u);
}
__device__ __forceinline__ float ubm(unsigned w, int k) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 + k)) - 8388608.0f; }

template <int ST> __device__ __forceinline__ void st1(float* p, float v) {
    if (ST == 0) __stcs(p, v); else if (ST == 1) *p = v; else __stwt(p, v);
}
template <int ST> __device__ __forceinline__ void st2(float* p, float2 v) {
    if (ST == 0) __stcs((float2*)p, v); else if (ST == 1) *(float2*)p = v; else __stwt((float2*)p, v);
}
template <int ST> __device__ __forceinline__ void st4(float* p, float4 v) {
    if (ST == 0) __stcs((float4*)p, v); else if (ST == 1) *(float4*)p = v; else __stwt((float4*)p, v);
}

// one gather sample of pixel (x, y): uint8 2x2 entry, one LDG.128
__device__ __forceinline__ void sample(const uint4* __restrict__ uq, const float* __restrict__ fl, int b, int x, int y, int n, int f, float (&acc)[3]) {
    const long long p = (long long)y * W + x;
    const float* F = fl + (long long)b * 4 * NPX + p;
    const float t = (n + 1) * 0.125f;
    const float c0 = f ? (1 - t) * (1 - t) : -(1 - t) * t, c1 = f ? -t * (1 - t) : t * t;
    const float u = c0 * __ldg(F) + c1 * __ldg(F + 2 * NPX), v = c0 * __ldg(F + NPX) + c1 * __ldg(F + 3 * NPX);
    const float ix = fminf(fmaxf(x + u, 0.f), W - 1.001f), iy = fminf(fmaxf(y + v, 0.f), H - 1.001f);
    const int x0 = (int)ix, y0 = (int)iy; const float wx = ix - x0, wy = iy - y0;
    const uint4 q = __ldg(uq + ((long long)b * 2 + f) * NPX + (long long)y0 * W + x0);
    const unsigned w[3] = {q.x, q.y, q.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const int byte = 3 * j + k; v4[j] = ubm(w[byte >> 2], byte & 3); }
        acc[k] += (v4[0] * (1 - wx) + v4[1] * wx) * (1 - wy) + (v4[2] * (1 - wx) + v4[3] * wx) * wy;
    }
}

// TW x TH pixel tile per 256-thread CTA, PPT consecutive pixels per thread (TW = 32 * PPT * warps_per_row)
template <int TW, int TH, int PPT, int ST, bool GATHER>
__global__ void __launch_bounds__(256, 4) k_planar(const uint4* __restrict__ uq, const float* __restrict__ fl, float* __restrict__ out) {
    constexpr int TPR = TW / PPT;                       // threads per tile row
    static_assert(TPR * TH == 256, "256 threads");
    const int tiles_x = W / TW, tiles_y = H / TH, tpp = tiles_x * tiles_y;
    const int b = blockIdx.x / tpp, r = blockIdx.x - b * tpp, ty = r / tiles_x, tx = r - ty * tiles_x;
    const int x = tx * TW + (threadIdx.x % TPR) * PPT, y = ty * TH + threadIdx.x / TPR;
    const long long p = (long long)y * W + x;
    for (int n = 0; n < N; ++n) {
        float v[PPT][3];
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            v[i][0] = (float)(x + i); v[i][1] = (float)y; v[i][2] = (float)n;
            if (GATHER) { sample(uq, fl, b, x + i, y, n, 0, v[i]); sample(uq, fl, b, x + i, y, n, 1, v[i]); }
        }
        float* O = out + ((long long)(b * N + n) * C) * NPX + p;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            float* o = O + (long long)k * NPX;
            if (PPT == 1) st1<ST>(o, v[0][k % 3] + k);
            else if (PPT == 2) st2<ST>(o, make_float2(v[0][k % 3] + k, v[1][k % 3] + k));
            else st4<ST>(o, make_float4(v[0][k % 3] + k, v[1][k % 3] + k, v[2][k % 3] + k, v[3][k % 3] + k));
        }
    }
}
__device__ __forceinline__ void stcs256(void* p, const float (&w)[8]) {
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(w[0]), "f"(w[1]), "f"(w[2]), "f"(w[3]), "f"(w[4]), "f"(w[5]), "f"(w[6]), "f"(w[7]) : "memory");
}
template <bool GATHER>
__global__ void __launch_bounds__(256, 4) k_nhwc(const uint4* __restrict__ uq, const float* __restrict__ fl, float* __restrict__ out) {
    const int tiles_x = W / 32, tiles_y = H / 8, tpp = tiles_x * tiles_y;
    const int b = blockIdx.x / tpp, r = blockIdx.x - b * tpp, ty = r / tiles_x, tx = r - ty * tiles_x;
    const int x = tx * 32 + (threadIdx.x & 31), y = ty * 8 + (threadIdx.x >> 5);
    const long long p = (long long)y * W + x;
    for (int n = 0; n < N; ++n) {
        float v[3] = {(float)x, (float)y, (float)n};
        if (GATHER) { sample(uq, fl, b, x, y, n, 0, v); sample(uq, fl, b, x, y, n, 1, v); }
        float* O = out + ((long long)(b * N + n) * NPX + p) * C;
        float a[8], c[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { a[k] = v[k % 3] + k; c[k] = v[(k + 8) % 3] + k + 8; }
        stcs256(O, a); stcs256(O + 8, c);
    }
}
__global__ void k_fill(float4* out, long long n4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        __stcs(out + i, make_float4(1.f, 2.f, 3.f, (float)i));
}

template <typename F> float time_ms(F&& launch, int reps = 10) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms / reps;
}

int main() {
    float *flow, *big; uint4* uq;
    const size_t out_bytes = (size_t)B * N * C * NPX * 4;
    CK(cudaMalloc(&flow, B * 4 * NPX * 4)); CK(cudaMalloc(&uq, (size_t)B * 2 * NPX * 16)); CK(cudaMalloc(&big, out_bytes));
    const int T = 256;
    const long long nf = (long long)B * 4 * NPX, nq = (long long)B * 2 * NPX;
    make_flow<<<(unsigned)((nf + T - 1) / T), T>>>(flow, 8, 20.0f, 12345u, nf);
    make_quads<<<(unsigned)((nq + T - 1) / T), T>>>(uq, nq);
    CK(cudaDeviceSynchronize());
    struct R { const char* name; float ms; };
    std::vector<R> rs;
    auto run = [&](const char* name, auto&& fn) { rs.push_back({name, time_ms(fn)}); };
    run("fill_stcs_v4", [&] { k_fill<<<148 * 16, 256>>>((float4*)big, (long long)(out_bytes / 16)); });
#define SHAPE(TW, TH, PPT, ST, G, NAME) run(NAME, [&] { k_planar<TW, TH, PPT, ST, G><<<B * (W / TW) * (H / TH), 256>>>(uq, flow, big); });
    SHAPE(32, 8, 1, 0, false, "store_32x8_p1_cs") SHAPE(32, 8, 1, 1, false, "store_32x8_p1_default") SHAPE(32, 8, 1, 2, false, "store_32x8_p1_wt")
    SHAPE(64, 4, 1, 0, false, "store_64x4_p1_cs") SHAPE(128, 2, 1, 0, false, "store_128x2_p1_cs")
    SHAPE(64, 8, 2, 0, false, "store_64x8_p2_cs") SHAPE(128, 8, 4, 0, false, "store_128x8_p4_cs") SHAPE(128, 8, 4, 1, false, "store_128x8_p4_default")
    SHAPE(128, 4, 2, 0, false, "store_128x4_p2_cs")
    run("store_nhwc", [&] { k_nhwc<false><<<B * (W / 32) * (H / 8), 256>>>(uq, flow, big); });
    SHAPE(32, 8, 1, 0, true, "k1u8_32x8_p1_cs") SHAPE(32, 8, 1, 1, true, "k1u8_32x8_p1_default") SHAPE(32, 8, 1, 2, true, "k1u8_32x8_p1_wt")
    SHAPE(64, 4, 1, 0, true, "k1u8_64x4_p1_cs") SHAPE(128, 2, 1, 0, true, "k1u8_128x2_p1_cs")
    SHAPE(64, 8, 2, 0, true, "k1u8_64x8_p2_cs") SHAPE(128, 8, 4, 0, true, "k1u8_128x8_p4_cs") SHAPE(128, 4, 2, 0, true, "k1u8_128x4_p2_cs")
    run("k1u8_nhwc", [&] { k_nhwc<true><<<B * (W / 32) * (H / 8), 256>>>(uq, flow, big); });
    for (auto& r : rs)
        printf("{\"variant\": \"%s\", \"ms\": %.3f, \"store_tbs\": %.2f}\n", r.name, r.ms, out_bytes / (r.ms * 1e-3) / 1e12);
    return 0;
}
