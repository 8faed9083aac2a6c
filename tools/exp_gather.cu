// exp_gather.cu -- GPU experiment (not product code, not a bench number): what does one bilinear
// 2x2x3-channel gather cost on sm_100a as a function of the frame layout and the load path?
//
// The forward kernels are bound by the L1 data pipe (profiles/r01b: 89 % / 81 % l1tex data-pipe
// wavefronts, ~11.7 wavefronts per 16-byte-per-lane gather request).  This program times gather-only
// kernels (7 timesteps x 2 frames x 4 taps x 3 channels per pixel, one float written per pixel) on a
// 16-pair 1088x1920 workload with the bench's rough flow (control grid at 1/8 resolution) and a
// smooth one (1/64), for these variants:
//   planar4    planar fp32 frames, 12 LDG.32 per sample              (round 1a kernels)
//   rgbx16     RGBx 16 B texels, 4 LDG.128 per sample                (round 1b kernels)
//   pair32     pair-packed 32 B entries {texel x, texel x+1}, 2 LDG.256 per sample
//   rgbx16_b84 rgbx16 with a warp covering an 8x4 pixel block instead of 32x1
//   texgather  layered R32F cudaArray, 3 tex2Dgather per sample
//   texpoint   layered RGBA32F cudaArray, 4 point tex2DLayered<float4> per sample
//   mixed      frame 0 through rgbx16, frame 1 through texgather
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/exp_gather tools/exp_gather.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int H = 1088, W = 1920, B = 16, N = 7;
constexpr long long NPX = (long long)H * W;

__device__ __forceinline__ unsigned hash_u(unsigned a) {
    a ^= a >> 16; a *= 0x7feb352dU; a ^= a >> 15; a *= 0x846ca68bU; a ^= a >> 16; return a;
}
__device__ __forceinline__ float randn_(unsigned k) {   // Box-Muller on two hashes
    float u1 = (hash_u(k * 2 + 1) >> 8) * (1.0f / 16777216.0f) + 1e-7f;
    float u2 = (hash_u(k * 2 + 2) >> 8) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
// control grid of normal deviates at 1/G resolution, bilinearly upsampled (align_corners=False)
__global__ void make_flow(float* flow, int G, float px, unsigned seed) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)B * 4 * NPX) return;
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
__global__ void make_img(float* img) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)B * 6 * NPX) return;
    img[i] = (hash_u((unsigned)i) >> 8) * (1.0f / 16777216.0f);
}
__global__ void pack16(const float* img, float4* out) {   // B x 2 x H x W texels
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)B * 2 * NPX) return;
    long long p = i % NPX; long long bf = i / NPX;
    const float* s = img + bf * 3 * NPX + p;
    out[i] = make_float4(s[0], s[NPX], s[2 * NPX], 0.f);
}
__global__ void pack32(const float* img, float4* out) {   // entry x: texel x, texel x+1
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)B * 2 * NPX) return;
    long long p = i % NPX; long long bf = i / NPX; int x = p % W;
    const float* s = img + bf * 3 * NPX + p;
    int d = (x + 1 < W) ? 1 : 0;
    out[2 * i] = make_float4(s[0], s[NPX], s[2 * NPX], d ? s[1] : 0.f);
    out[2 * i + 1] = make_float4(d ? s[NPX + 1] : 0.f, d ? s[2 * NPX + 1] : 0.f, 0.f, 0.f);
}

struct Smp { int x0, y0; float wx, wy; };
// position of the sample of pixel (x,y), timestep n, frame f (clamped inside so no masks are needed)
__device__ __forceinline__ Smp sample_pos(const float* __restrict__ fl, int b, int x, int y, int n, int f) {
    long long p = (long long)y * W + x;
    const float* F = fl + (long long)b * 4 * NPX + p;
    float t = (n + 1) * 0.125f;
    float c0 = f ? (1 - t) * (1 - t) : -(1 - t) * t, c1 = f ? -t * (1 - t) : t * t;
    float u = c0 * __ldg(F + (f ? 0 : 0)) + c1 * __ldg(F + 2 * NPX);
    float v = c0 * __ldg(F + NPX) + c1 * __ldg(F + 3 * NPX);
    float ix = fminf(fmaxf(x + u, 0.f), W - 1.001f), iy = fminf(fmaxf(y + v, 0.f), H - 1.001f);
    Smp s; s.x0 = (int)ix; s.y0 = (int)iy; s.wx = ix - s.x0; s.wy = iy - s.y0;
    return s;
}
template <int SHAPE> __device__ __forceinline__ bool pixel_of_thread(int& b, int& x, int& y) {
    // SHAPE 0: CTA = 32x8 tile, warp = 32x1 row.  SHAPE 1: CTA = 32x8 tile, warp = 8x4 block.
    int tiles_x = W / 32, tiles_y = H / 8, tpp = tiles_x * tiles_y;
    b = blockIdx.x / tpp; int r = blockIdx.x - b * tpp; int ty = r / tiles_x, tx = r - ty * tiles_x;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (SHAPE == 0) { x = tx * 32 + lane; y = ty * 8 + wid; }
    else { x = tx * 32 + (wid & 3) * 8 + (lane & 7); y = ty * 8 + (wid >> 2) * 4 + (lane >> 3); }
    return true;
}
__device__ __forceinline__ float lerp4(float a, float b, float c, float d, float wx, float wy) {
    return (a * (1 - wx) + b * wx) * (1 - wy) + (c * (1 - wx) + d * wx) * wy;
}

template <int SHAPE>
__global__ void __launch_bounds__(256) k_planar4(const float* __restrict__ img, const float* __restrict__ fl, float* __restrict__ out) {
    int b, x, y; pixel_of_thread<SHAPE>(b, x, y);
    float acc = 0.f;
    for (int n = 0; n < N; ++n)
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            Smp s = sample_pos(fl, b, x, y, n, f);
            const float* q = img + ((long long)b * 6 + f * 3) * NPX + (long long)s.y0 * W + s.x0;
#pragma unroll
            for (int c = 0; c < 3; ++c, q += NPX) acc += lerp4(__ldg(q), __ldg(q + 1), __ldg(q + W), __ldg(q + W + 1), s.wx, s.wy);
        }
    out[(long long)b * NPX + (long long)y * W + x] = acc;
}
template <int SHAPE>
__global__ void __launch_bounds__(256) k_rgbx16(const float4* __restrict__ img, const float* __restrict__ fl, float* __restrict__ out) {
    int b, x, y; pixel_of_thread<SHAPE>(b, x, y);
    float acc = 0.f;
    for (int n = 0; n < N; ++n)
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            Smp s = sample_pos(fl, b, x, y, n, f);
            const float4* q = img + ((long long)b * 2 + f) * NPX + (long long)s.y0 * W + s.x0;
            float4 a = __ldg(q), bb = __ldg(q + 1), c = __ldg(q + W), d = __ldg(q + W + 1);
            acc += lerp4(a.x, bb.x, c.x, d.x, s.wx, s.wy) + lerp4(a.y, bb.y, c.y, d.y, s.wx, s.wy) + lerp4(a.z, bb.z, c.z, d.z, s.wx, s.wy);
        }
    out[(long long)b * NPX + (long long)y * W + x] = acc;
}
__device__ __forceinline__ void ld256(const float4* p, float (&a)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]) : "l"(p));
}
template <int SHAPE>
__global__ void __launch_bounds__(256) k_pair32(const float4* __restrict__ img, const float* __restrict__ fl, float* __restrict__ out) {
    int b, x, y; pixel_of_thread<SHAPE>(b, x, y);
    float acc = 0.f;
    for (int n = 0; n < N; ++n)
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            Smp s = sample_pos(fl, b, x, y, n, f);
            const float4* q = img + 2 * (((long long)b * 2 + f) * NPX + (long long)s.y0 * W + s.x0);
            float r0[8], r1[8];
            ld256(q, r0); ld256(q + 2 * W, r1);
            acc += lerp4(r0[0], r0[3], r1[0], r1[3], s.wx, s.wy) + lerp4(r0[1], r0[4], r1[1], r1[4], s.wx, s.wy) + lerp4(r0[2], r0[5], r1[2], r1[5], s.wx, s.wy);
        }
    out[(long long)b * NPX + (long long)y * W + x] = acc;
}
__device__ __forceinline__ float4 gather_layer(cudaTextureObject_t tex, float x, float y, int layer) {
    float4 r;   // no CUDA C intrinsic for a layered gather: tld4 on the a2d geometry
    asm volatile("{ .reg .f32 d; mov.f32 d, 0f00000000;\n\t"
                 "tld4.r.a2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6,%7,d}]; }"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(tex), "r"(layer), "f"(x), "f"(y));
    return r;
}
// tex variants: texture objects over layered arrays, layer = b*6 + f*3 + c (R32F) or b*2 + f (RGBA32F)
template <int SHAPE, bool MIXED>
__global__ void __launch_bounds__(256) k_texgather(cudaTextureObject_t tex, const float4* __restrict__ img, const float* __restrict__ fl, float* __restrict__ out) {
    int b, x, y; pixel_of_thread<SHAPE>(b, x, y);
    float acc = 0.f;
    for (int n = 0; n < N; ++n)
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            Smp s = sample_pos(fl, b, x, y, n, f);
            if (MIXED && f == 0) {
                const float4* q = img + ((long long)b * 2 + f) * NPX + (long long)s.y0 * W + s.x0;
                float4 a = __ldg(q), bb = __ldg(q + 1), c = __ldg(q + W), d = __ldg(q + W + 1);
                acc += lerp4(a.x, bb.x, c.x, d.x, s.wx, s.wy) + lerp4(a.y, bb.y, c.y, d.y, s.wx, s.wy) + lerp4(a.z, bb.z, c.z, d.z, s.wx, s.wy);
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    // gather returns (x0,y1) (x1,y1) (x1,y0) (x0,y0) as .x .y .z .w
                    const int l = b * 6 + f * 3 + c;   // layers tiled 4 across in one 2D array (gather is 2D-only)
                    float4 g = tex2Dgather<float4>(tex, (l & 3) * W + s.x0 + 1.0f, (l >> 2) * H + s.y0 + 1.0f, 0);
                    acc += lerp4(g.w, g.z, g.x, g.y, s.wx, s.wy);
                }
            }
        }
    out[(long long)b * NPX + (long long)y * W + x] = acc;
}
template <int SHAPE>
__global__ void __launch_bounds__(256) k_texpoint(cudaTextureObject_t tex, const float* __restrict__ fl, float* __restrict__ out) {
    int b, x, y; pixel_of_thread<SHAPE>(b, x, y);
    float acc = 0.f;
    for (int n = 0; n < N; ++n)
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            Smp s = sample_pos(fl, b, x, y, n, f);
            float fx = s.x0 + 0.5f, fy = s.y0 + 0.5f; int l = b * 2 + f;
            float4 a = tex2DLayered<float4>(tex, fx, fy, l), bb = tex2DLayered<float4>(tex, fx + 1, fy, l);
            float4 c = tex2DLayered<float4>(tex, fx, fy + 1, l), d = tex2DLayered<float4>(tex, fx + 1, fy + 1, l);
            acc += lerp4(a.x, bb.x, c.x, d.x, s.wx, s.wy) + lerp4(a.y, bb.y, c.y, d.y, s.wx, s.wy) + lerp4(a.z, bb.z, c.z, d.z, s.wx, s.wy);
        }
    out[(long long)b * NPX + (long long)y * W + x] = acc;
}
// reference of the non-gather part: the same kernel with the loads replaced by arithmetic
template <int SHAPE>
__global__ void __launch_bounds__(256) k_nogather(const float* __restrict__ fl, float* __restrict__ out) {
    int b, x, y; pixel_of_thread<SHAPE>(b, x, y);
    float acc = 0.f;
    for (int n = 0; n < N; ++n)
#pragma unroll
        for (int f = 0; f < 2; ++f) { Smp s = sample_pos(fl, b, x, y, n, f); acc += s.wx * s.wy + s.x0 + s.y0; }
    out[(long long)b * NPX + (long long)y * W + x] = acc;
}

// K1-like / K2-like kernels: the gathers of both frames per timestep PLUS the streaming traffic of
// the real kernels (K2: 5 loads + 3 stores per timestep with a +-0.5 px white-noise residual on the
// sample position; K1: 16 stores per timestep), with TEXW of the 8 warps of a CTA gathering through
// the texture unit (tex2Dgather on R32F) and the others through the LSU (RGBx LDG.128).
template <int TEXW, int NSTORE, bool JITTER>
__global__ void __launch_bounds__(256, 4) k_like(cudaTextureObject_t tex, const float4* __restrict__ img, const float* __restrict__ fl,
                                                 const float* __restrict__ y5, float* __restrict__ out) {
    int b, x, y; pixel_of_thread<0>(b, x, y);
    const bool use_tex = (threadIdx.x >> 5) < TEXW;
    const long long p = (long long)y * W + x;
    for (int n = 0; n < N; ++n) {
        float jx0 = 0, jy0 = 0, jx1 = 0, jy1 = 0, lg = 0;
        if (JITTER) {
            const float* Y = y5 + ((long long)(b * N + n) * 5) * NPX + p;
            lg = __ldcs(Y); jx1 = __ldcs(Y + NPX); jy1 = __ldcs(Y + 2 * NPX); jx0 = __ldcs(Y + 3 * NPX); jy0 = __ldcs(Y + 4 * NPX);
        }
        float acc[3] = {lg, 0.f, 0.f};
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            Smp s = sample_pos(fl, b, x, y, n, f);
            if (JITTER) {
                float ix = fminf(fmaxf(s.x0 + s.wx + (f ? jx1 : jx0), 0.f), W - 1.001f), iy = fminf(fmaxf(s.y0 + s.wy + (f ? jy1 : jy0), 0.f), H - 1.001f);
                s.x0 = (int)ix; s.y0 = (int)iy; s.wx = ix - s.x0; s.wy = iy - s.y0;
            }
            if (use_tex) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int l = b * 6 + f * 3 + c;
                    float4 g = tex2Dgather<float4>(tex, (l & 3) * W + s.x0 + 1.0f, (l >> 2) * H + s.y0 + 1.0f, 0);
                    acc[c] += lerp4(g.w, g.z, g.x, g.y, s.wx, s.wy);
                }
            } else {
                const float4* q = img + ((long long)b * 2 + f) * NPX + (long long)s.y0 * W + s.x0;
                float4 a = __ldg(q), bb = __ldg(q + 1), c = __ldg(q + W), d = __ldg(q + W + 1);
                acc[0] += lerp4(a.x, bb.x, c.x, d.x, s.wx, s.wy); acc[1] += lerp4(a.y, bb.y, c.y, d.y, s.wx, s.wy); acc[2] += lerp4(a.z, bb.z, c.z, d.z, s.wx, s.wy);
            }
        }
        float* O = out + ((long long)(b * N + n) * NSTORE) * NPX + p;
#pragma unroll
        for (int k = 0; k < NSTORE; ++k) __stcs(O + (long long)k * NPX, acc[k % 3] + k);
    }
}


// lane-paired gathers: lanes (2k, 2k+1) work on the two pixels of the pair together.  Each LDG.128 then
// covers 16 samples x 32 contiguous bytes (west texel in the even lane, east texel in the odd lane) instead
// of 32 samples x 16 bytes, so a request touches about half as many distinct rows.  Each lane computes the
// west (or east) half of the bilinear sum of BOTH pixels and the halves are exchanged with SHFL.
struct LP { float4 a0, a1, b0, b1; float wxA, wyA, wxB, wyB; };
__device__ __forceinline__ void lp_gather(const float4* __restrict__ plane, const Smp& s, int role, float (&res)[3]) {
    int off = s.y0 * W + s.x0;
    int offp = __shfl_xor_sync(0xffffffffu, off, 1);
    float wxp = __shfl_xor_sync(0xffffffffu, s.wx, 1), wyp = __shfl_xor_sync(0xffffffffu, s.wy, 1);
    const int offA = role ? offp : off, offB = role ? off : offp;
    const float wxA = role ? wxp : s.wx, wyA = role ? wyp : s.wy, wxB = role ? s.wx : wxp, wyB = role ? s.wy : wyp;
    const float4* q = plane + role;
    float4 a0 = __ldg(q + offA), a1 = __ldg(q + offA + W), b0 = __ldg(q + offB), b1 = __ldg(q + offB + W);
    const float hA = role ? wxA : 1 - wxA, hB = role ? wxB : 1 - wxB;
    float pA[3] = {hA * (a0.x * (1 - wyA) + a1.x * wyA), hA * (a0.y * (1 - wyA) + a1.y * wyA), hA * (a0.z * (1 - wyA) + a1.z * wyA)};
    float pB[3] = {hB * (b0.x * (1 - wyB) + b1.x * wyB), hB * (b0.y * (1 - wyB) + b1.y * wyB), hB * (b0.z * (1 - wyB) + b1.z * wyB)};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float got = __shfl_xor_sync(0xffffffffu, role ? pA[c] : pB[c], 1);
        res[c] = (role ? pB[c] : pA[c]) + got;
    }
}
__global__ void __launch_bounds__(256) k_lanepair(const float4* __restrict__ img, const float* __restrict__ fl, float* __restrict__ out) {
    int b, x, y; pixel_of_thread<0>(b, x, y);
    const int role = threadIdx.x & 1;
    float acc = 0.f;
    for (int n = 0; n < N; ++n)
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            Smp s = sample_pos(fl, b, x, y, n, f);
            float r[3]; lp_gather(img + ((long long)b * 2 + f) * NPX, s, role, r);
            acc += r[0] + r[1] + r[2];
        }
    out[(long long)b * NPX + (long long)y * W + x] = acc;
}
template <int NSTORE, bool JITTER>
__global__ void __launch_bounds__(256, 4) k_like_lp(const float4* __restrict__ img, const float* __restrict__ fl,
                                                    const float* __restrict__ y5, float* __restrict__ out) {
    int b, x, y; pixel_of_thread<0>(b, x, y);
    const int role = threadIdx.x & 1;
    const long long p = (long long)y * W + x;
    for (int n = 0; n < N; ++n) {
        float jx0 = 0, jy0 = 0, jx1 = 0, jy1 = 0, lg = 0;
        if (JITTER) {
            const float* Y = y5 + ((long long)(b * N + n) * 5) * NPX + p;
            lg = __ldcs(Y); jx1 = __ldcs(Y + NPX); jy1 = __ldcs(Y + 2 * NPX); jx0 = __ldcs(Y + 3 * NPX); jy0 = __ldcs(Y + 4 * NPX);
        }
        float acc[3] = {lg, 0.f, 0.f};
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            Smp s = sample_pos(fl, b, x, y, n, f);
            if (JITTER) {
                float ix = fminf(fmaxf(s.x0 + s.wx + (f ? jx1 : jx0), 0.f), W - 1.001f), iy = fminf(fmaxf(s.y0 + s.wy + (f ? jy1 : jy0), 0.f), H - 1.001f);
                s.x0 = (int)ix; s.y0 = (int)iy; s.wx = ix - s.x0; s.wy = iy - s.y0;
            }
            float r[3]; lp_gather(img + ((long long)b * 2 + f) * NPX, s, role, r);
            acc[0] += r[0]; acc[1] += r[1]; acc[2] += r[2];
        }
        float* O = out + ((long long)(b * N + n) * NSTORE) * NPX + p;
#pragma unroll
        for (int k = 0; k < NSTORE; ++k) __stcs(O + (long long)k * NPX, acc[k % 3] + k);
    }
}

// Which model describes the L1 data stage?  MODE 0: the 8 lanes of a quarter-warp read the same column of 8 different
// rows (8 lines, one 16-byte bank group).  MODE 1: 8 different rows AND 8 different columns mod 8 (8 lines, 8 bank
// groups).  MODE 2: 8 consecutive texels of one row (1 line).  A "one wavefront per line" model predicts 8/8/1
// wavefronts per quarter-warp, a banked-SRAM model 8/1/1.
template <int MODE>
__global__ void __launch_bounds__(256) k_bank(const float4* __restrict__ img, float* __restrict__ out) {
    int b, x, y; pixel_of_thread<0>(b, x, y);
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int n = 0; n < 14; ++n) {
        int yy = (y + 37 * n) % (H - 40), xx = x - lane;
        int row = MODE == 2 ? yy : yy + lane, col = MODE == 0 ? xx + (lane >> 3) : xx + lane;
        const float4* q = img + (long long)b * 2 * NPX + (long long)row * W + col;
        float4 a = __ldg(q), c = __ldg(q + 2 * W), d = __ldg(q + 4 * W), e = __ldg(q + 6 * W);
        acc += a.x + c.y + d.z + e.x;
    }
    out[(long long)b * NPX + (long long)y * W + x] = acc;
}
// rgbx16 with the taps read around L1 (ld.global.cg): does a miss path cost the same data-stage wavefronts?
__global__ void __launch_bounds__(256) k_rgbx16cg(const float4* __restrict__ img, const float* __restrict__ fl, float* __restrict__ out) {
    int b, x, y; pixel_of_thread<0>(b, x, y);
    float acc = 0.f;
    for (int n = 0; n < N; ++n)
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            Smp s = sample_pos(fl, b, x, y, n, f);
            const float4* q = img + ((long long)b * 2 + f) * NPX + (long long)s.y0 * W + s.x0;
            float4 a = __ldcg(q), bb = __ldcg(q + 1), c = __ldcg(q + W), d = __ldcg(q + W + 1);
            acc += lerp4(a.x, bb.x, c.x, d.x, s.wx, s.wy) + lerp4(a.y, bb.y, c.y, d.y, s.wx, s.wy) + lerp4(a.z, bb.z, c.z, d.z, s.wx, s.wy);
        }
    out[(long long)b * NPX + (long long)y * W + x] = acc;
}

// per-THREAD mix of the two load paths over the caller's own RGBx buffer: the first NT of the 8 taps of a timestep
// (2 frames x 4 corners) go through the texture unit as point fetches from a LINEAR-memory texture object
// (tex1Dfetch<float4>, no cudaArray, no upload), the others through the LSU (LDG.128).
template <int NT, int NSTORE, bool JITTER>
__global__ void __launch_bounds__(256, 4) k_like_mixlin(cudaTextureObject_t lin, const float4* __restrict__ img, const float* __restrict__ fl,
                                                        const float* __restrict__ y5, float* __restrict__ out) {
    int b, x, y; pixel_of_thread<0>(b, x, y);
    const long long p = (long long)y * W + x;
    for (int n = 0; n < N; ++n) {
        float jx0 = 0, jy0 = 0, jx1 = 0, jy1 = 0, lg = 0;
        if (JITTER) {
            const float* Y = y5 + ((long long)(b * N + n) * 5) * NPX + p;
            lg = __ldcs(Y); jx1 = __ldcs(Y + NPX); jy1 = __ldcs(Y + 2 * NPX); jx0 = __ldcs(Y + 3 * NPX); jy0 = __ldcs(Y + 4 * NPX);
        }
        float acc[3] = {lg, 0.f, 0.f};
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            Smp s = sample_pos(fl, b, x, y, n, f);
            if (JITTER) {
                float ix = fminf(fmaxf(s.x0 + s.wx + (f ? jx1 : jx0), 0.f), W - 1.001f), iy = fminf(fmaxf(s.y0 + s.wy + (f ? jy1 : jy0), 0.f), H - 1.001f);
                s.x0 = (int)ix; s.y0 = (int)iy; s.wx = ix - s.x0; s.wy = iy - s.y0;
            }
            const int base = (b * 2 + f) * (int)NPX + s.y0 * W + s.x0;
            const int off[4] = {0, 1, W, W + 1};
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (f * 4 + k < NT) v[k] = tex1Dfetch<float4>(lin, base + off[k]);
                else v[k] = __ldg(img + base + off[k]);
            }
            acc[0] += lerp4(v[0].x, v[1].x, v[2].x, v[3].x, s.wx, s.wy); acc[1] += lerp4(v[0].y, v[1].y, v[2].y, v[3].y, s.wx, s.wy);
            acc[2] += lerp4(v[0].z, v[1].z, v[2].z, v[3].z, s.wx, s.wy);
        }
        float* O = out + ((long long)(b * N + n) * NSTORE) * NPX + p;
#pragma unroll
        for (int k = 0; k < NSTORE; ++k) __stcs(O + (long long)k * NPX, acc[k % 3] + k);
    }
}

static bool g_once = false;   // `exp_gather once`: one launch per variant, rough flow only (for ncu counters)
template <typename F> float time_ms(F&& launch, int reps = 10) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    if (g_once) reps = 1;
    for (int i = 0; i < (g_once ? 0 : 3); ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms / reps;
}
static double checksum(const float* d_out) {
    std::vector<float> h(NPX);
    CK(cudaMemcpy(h.data(), d_out, NPX * 4, cudaMemcpyDeviceToHost));
    double s = 0; for (long long i = 0; i < NPX; ++i) s += h[i];
    return s / NPX;
}

int main(int argc, char** argv) {
    g_once = argc > 1 && !strcmp(argv[1], "once");
    float *img, *flow, *out; float4 *p16, *p32;
    CK(cudaMalloc(&img, B * 6 * NPX * 4)); CK(cudaMalloc(&flow, B * 4 * NPX * 4)); CK(cudaMalloc(&out, B * NPX * 4));
    CK(cudaMalloc(&p16, B * 2 * NPX * 16)); CK(cudaMalloc(&p32, B * 2 * NPX * 32));
    const int T = 256;
    make_img<<<(B * 6 * NPX + T - 1) / T, T>>>(img);
    pack16<<<(B * 2 * NPX + T - 1) / T, T>>>(img, p16);
    pack32<<<(B * 2 * NPX + T - 1) / T, T>>>(img, p32);
    CK(cudaDeviceSynchronize());

    // layered arrays
    cudaChannelFormatDesc d1 = cudaCreateChannelDesc<float>(), d4 = cudaCreateChannelDesc<float4>();
    cudaArray_t arr1, arr4;
    CK(cudaMallocArray(&arr1, &d1, W * 4, H * (B * 6 / 4), cudaArrayTextureGather));
    CK(cudaMalloc3DArray(&arr4, &d4, make_cudaExtent(W, H, B * 2), cudaArrayLayered));
    for (int l = 0; l < B * 6; ++l)
        CK(cudaMemcpy2DToArray(arr1, (l & 3) * W * 4, (l >> 2) * H, img + l * NPX, W * 4, W * 4, H, cudaMemcpyDeviceToDevice));
    cudaMemcpy3DParms cq = {};
    cq.srcPtr = make_cudaPitchedPtr(p16, W * 16, W, H); cq.dstArray = arr4; cq.extent = make_cudaExtent(W, H, B * 2); cq.kind = cudaMemcpyDeviceToDevice;
    CK(cudaMemcpy3D(&cq));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray;
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder; td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t tex1, tex4;
    rd.res.array.array = arr1; CK(cudaCreateTextureObject(&tex1, &rd, &td, nullptr));
    rd.res.array.array = arr4; CK(cudaCreateTextureObject(&tex4, &rd, &td, nullptr));

    cudaTextureObject_t texlin;
    {
        cudaResourceDesc rl = {}; rl.resType = cudaResourceTypeLinear; rl.res.linear.devPtr = p16; rl.res.linear.desc = d4;
        rl.res.linear.sizeInBytes = (size_t)B * 2 * NPX * 16;
        cudaTextureDesc tl = {}; tl.addressMode[0] = cudaAddressModeClamp; tl.filterMode = cudaFilterModePoint; tl.readMode = cudaReadModeElementType;
        CK(cudaCreateTextureObject(&texlin, &rl, &tl, nullptr));
    }
    // time the array upload paths too (what a pack pre-pass into an array would cost at best)
    float up1 = time_ms([&] { for (int l = 0; l < B * 6; ++l) CK(cudaMemcpy2DToArrayAsync(arr1, (l & 3) * W * 4, (l >> 2) * H, img + l * NPX, W * 4, W * 4, H, cudaMemcpyDeviceToDevice)); }, 3), up4 = time_ms([&] { CK(cudaMemcpy3DAsync(&cq)); }, 3);
    printf("{\"upload_r32f_layers_ms\": %.3f, \"upload_rgba32f_layers_ms\": %.3f}\n", up1, up4);

    const int grid = B * (W / 32) * (H / 8);
    float *y5, *big;
    CK(cudaMalloc(&y5, (size_t)B * N * 5 * NPX * 4)); CK(cudaMalloc(&big, (size_t)B * N * 16 * NPX * 4));
    make_flow<<<(int)(((long long)B * 4 * NPX + T - 1) / T), T>>>(y5, 1, 0.5f, 777u);      // white noise, sigma 0.5 px (first 4/35 of y5 ...)
    for (int r = 0; r < 9; ++r) make_flow<<<(int)(((long long)B * 4 * NPX + T - 1) / T), T>>>(y5 + (size_t)r * B * 4 * NPX * 7 / 8, 1, 0.5f, 778u + r);
    CK(cudaDeviceSynchronize());
    const double samples = (double)B * NPX * N * 2;
    const int grids[2] = {8, 64};
    for (int gi = 0; gi < (g_once ? 1 : 2); ++gi) {
        make_flow<<<(B * 4 * NPX + T - 1) / T, T>>>(flow, grids[gi], 20.0f, 12345u);
        CK(cudaDeviceSynchronize());
        struct R { const char* name; float ms; double sum; };
        std::vector<R> rs;
        auto run = [&](const char* name, auto&& fn) { float ms = time_ms(fn); rs.push_back({name, ms, checksum(out)}); };
        run("nogather", [&] { k_nogather<0><<<grid, 256>>>(flow, out); });
        run("planar4", [&] { k_planar4<0><<<grid, 256>>>(img, flow, out); });
        run("rgbx16", [&] { k_rgbx16<0><<<grid, 256>>>(p16, flow, out); });
        run("bank_samecol", [&] { k_bank<0><<<grid, 256>>>(p16, out); });
        run("bank_diag", [&] { k_bank<1><<<grid, 256>>>(p16, out); });
        run("bank_row", [&] { k_bank<2><<<grid, 256>>>(p16, out); });
        run("rgbx16cg", [&] { k_rgbx16cg<<<grid, 256>>>(p16, flow, out); });
        run("lanepair", [&] { k_lanepair<<<grid, 256>>>(p16, flow, out); });
        run("pair32", [&] { k_pair32<0><<<grid, 256>>>(p32, flow, out); });
        run("planar4_b84", [&] { k_planar4<1><<<grid, 256>>>(img, flow, out); });
        run("rgbx16_b84", [&] { k_rgbx16<1><<<grid, 256>>>(p16, flow, out); });
        run("pair32_b84", [&] { k_pair32<1><<<grid, 256>>>(p32, flow, out); });
        run("texgather", [&] { k_texgather<0, false><<<grid, 256>>>(tex1, p16, flow, out); });
        run("texgather_b84", [&] { k_texgather<1, false><<<grid, 256>>>(tex1, p16, flow, out); });
        run("texpoint", [&] { k_texpoint<0><<<grid, 256>>>(tex4, flow, out); });
        run("texpoint_b84", [&] { k_texpoint<1><<<grid, 256>>>(tex4, flow, out); });
        run("mixed", [&] { k_texgather<0, true><<<grid, 256>>>(tex1, p16, flow, out); });
        run("mixed_b84", [&] { k_texgather<1, true><<<grid, 256>>>(tex1, p16, flow, out); });
#define LIKE(TW) \
        run("k2like_texw" #TW, [&] { k_like<TW, 3, true><<<grid, 256>>>(tex1, p16, flow, y5, big); }); \
        run("k1like_texw" #TW, [&] { k_like<TW, 16, false><<<grid, 256>>>(tex1, p16, flow, y5, big); });
        run("k2like_lanepair", [&] { k_like_lp<3, true><<<grid, 256>>>(p16, flow, y5, big); });
        run("k1like_lanepair", [&] { k_like_lp<16, false><<<grid, 256>>>(p16, flow, y5, big); });
#define MIXLIN(NT) \
        run("k2like_mixlin" #NT, [&] { k_like_mixlin<NT, 3, true><<<grid, 256>>>(texlin, p16, flow, y5, big); }); \
        run("k1like_mixlin" #NT, [&] { k_like_mixlin<NT, 16, false><<<grid, 256>>>(texlin, p16, flow, y5, big); });
        MIXLIN(0) MIXLIN(1) MIXLIN(2) MIXLIN(3) MIXLIN(4) MIXLIN(8)
        LIKE(0) LIKE(2) LIKE(3) LIKE(4) LIKE(5) LIKE(6) LIKE(8)
        for (auto& r : rs)
            printf("{\"flow_grid\": \"1/%d\", \"variant\": \"%s\", \"ms\": %.3f, \"ns_per_warp_sample\": %.2f, \"cyc_per_warp_sample_per_sm\": %.1f, \"checksum\": %.6f}\n",
                   grids[gi], r.name, r.ms, r.ms * 1e6 / (samples / 32), r.ms * 1e-3 * 1.9e9 * 148 / (samples / 32), r.sum);
    }
    return 0;
}
