// ssm_abi.cu -- the C ABI of libssm_b200.so (declared in include/ssm_b200.h): argument checks,
// dtype / coordinate-mode dispatch and kernel launches.  CUDA only: there is no CPU path.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "ssm_frames.cuh"
#include "ssm_q8.cuh"
#include "ssm_scatter.cuh"
#include "ssm_unet_glue.cuh"

namespace {

using namespace ssm;

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e > 0 ? (int)e : 1;
}

#define SSM_LAUNCH_CHECK(what)                                   \
    do {                                                         \
        cudaError_t e__ = cudaGetLastError();                    \
        if (e__ != cudaSuccess) return cuda_fail(e__, what);     \
    } while (0)

constexpr size_t HDR_BYTES = 256;
constexpr long long MAX_PLANE = (1ll << 31) - (1 << 16);   // plane offsets are 32-bit

Geom make_geom(int H, int W) {
    Geom g;
    g.H = H; g.W = W;
    g.xnorm = (float)(W - 1 > 1 ? W - 1 : 1);
    g.ynorm = (float)(H - 1 > 1 ? H - 1 : 1);
    g.xinv = 1.0f / g.xnorm;
    g.yinv = 1.0f / g.ynorm;
    g.xm1 = (float)(W - 1);
    g.ym1 = (float)(H - 1);
    g.xgrad = g.xm1 / 2.0f;
    g.ygrad = g.ym1 / 2.0f;
    return g;
}

int check_common(int B, int N, int C, int H, int W, int dtype, int coord_mode) {
    if (B <= 0 || N <= 0 || C <= 0 || H <= 0 || W <= 0)
        return fail(SSM_ERR_SHAPE, "B, N, C, H, W must be positive (got B=%d N=%d C=%d H=%d W=%d)", B, N, C, H, W);
    if ((long long)H * W > MAX_PLANE) return fail(SSM_ERR_SHAPE, "H*W too large (%d x %d)", H, W);
    long long tiles = (long long)B * ((H + TILE_H - 1) / TILE_H) * ((W + TILE_W - 1) / TILE_W);
    if (tiles > 2147483647ll) return fail(SSM_ERR_SHAPE, "too many tiles for one launch (%lld)", tiles);
    if (dtype != SSM_DTYPE_F32 && dtype != SSM_DTYPE_BF16) return fail(SSM_ERR_DTYPE, "unknown dtype %d", dtype);
    if (coord_mode != SSM_COORD_DIV && coord_mode != SSM_COORD_RCP)
        return fail(SSM_ERR_DTYPE, "unknown coord_mode %d", coord_mode);
    return SSM_OK;
}

int check_tensor(const ssm_tensor* t, const char* name, int dtype, bool required) {
    if (t == nullptr || t->data == nullptr) {
        if (required) return fail(SSM_ERR_NULL, "%s is NULL", name);
        return SSM_OK;
    }
    size_t esz = dtype == SSM_DTYPE_F32 ? 4 : 2;
    if (((uintptr_t)t->data) % esz != 0) return fail(SSM_ERR_ALIGN, "%s is not aligned to its element size", name);
    const long long sc = t->stride_c < 0 ? -t->stride_c : t->stride_c;
    if (sc > ((1ll << 31) - 1) / 16) return fail(SSM_ERR_SHAPE, "%s: channel stride %lld too large (32-bit channel offsets)", name, sc);
    return SSM_OK;
}

int check_packed(const void* packed) {
    if (packed && ((uintptr_t)packed) % 16 != 0) return fail(SSM_ERR_ALIGN, "packed must be 16-byte aligned");
    return SSM_OK;
}

template <typename T> View<const T> cview(const ssm_tensor* t) {
    View<const T> v;
    if (t && t->data) { v.p = (const T*)t->data; v.sb = t->stride_b; v.sn = t->stride_n; v.sc = t->stride_c; }
    else { v.p = nullptr; v.sb = v.sn = v.sc = 0; }
    return v;
}
template <typename T> View<T> mview(const ssm_tensor* t) {
    View<T> v;
    if (t && t->data) { v.p = (T*)t->data; v.sb = t->stride_b; v.sn = t->stride_n; v.sc = t->stride_c; }
    else { v.p = nullptr; v.sb = v.sn = v.sc = 0; }
    return v;
}

unsigned tile_grid(int B, int H, int W) {
    return (unsigned)((long long)B * ((H + TILE_H - 1) / TILE_H) * ((W + TILE_W - 1) / TILE_W));
}

int count_bits_for(long long addends) {   // bits needed to hold `addends` unit contributions
    int b = 1;
    while ((1ll << b) <= addends) ++b;
    return b;
}

template <typename F> int dispatch3(int dtype, int mode, bool packed, F&& f) {
    if (dtype == SSM_DTYPE_F32) {
        if (mode == SSM_COORD_DIV) return packed ? f.template run<float, SSM_COORD_DIV, true>() : f.template run<float, SSM_COORD_DIV, false>();
        return packed ? f.template run<float, SSM_COORD_RCP, true>() : f.template run<float, SSM_COORD_RCP, false>();
    }
    if (mode == SSM_COORD_DIV)
        return packed ? f.template run<__nv_bfloat16, SSM_COORD_DIV, true>() : f.template run<__nv_bfloat16, SSM_COORD_DIV, false>();
    return packed ? f.template run<__nv_bfloat16, SSM_COORD_RCP, true>() : f.template run<__nv_bfloat16, SSM_COORD_RCP, false>();
}

// dispatch helper: calls f.template run<T, MODE>() for the runtime dtype / coord_mode
template <typename F> int dispatch(int dtype, int mode, F&& f) {
    if (dtype == SSM_DTYPE_F32)
        return mode == SSM_COORD_DIV ? f.template run<float, SSM_COORD_DIV>() : f.template run<float, SSM_COORD_RCP>();
    return mode == SSM_COORD_DIV ? f.template run<__nv_bfloat16, SSM_COORD_DIV>()
                                 : f.template run<__nv_bfloat16, SSM_COORD_RCP>();
}

unsigned sw_grid(int B, int H, int W) {      // one CTA per 64 x 16 tile of source pixels (ssm_scatter.cuh)
    return (unsigned)((long long)B * ((H + SW_TILE_H - 1) / SW_TILE_H) * ((W + SW_TILE_W - 1) / SW_TILE_W));
}

// scatter_finalize_kernel: planes x chunks of FIN_THREADS * FIN_PER_THREAD pixels (check_common bounds H*W and the tile count)
unsigned finalize_chunks(long long npx) { return (unsigned)((npx + FIN_THREADS * FIN_PER_THREAD - 1) / (FIN_THREADS * FIN_PER_THREAD)); }

// ---------------------------------------------------------------------------------------------
// packed != NULL: the image is the RGBx copy (B x H x W x 4, C = 3); img is then unused by the gathers
static ssm_tensor packed_image_desc(const void* packed, int H, int W) {
    ssm_tensor d;
    d.data = const_cast<void*>(packed); d.stride_b = (long long)H * W * 4; d.stride_n = 0; d.stride_c = 1;
    return d;
}

struct WarpFwd {
    const ssm_tensor *img, *flow, *out; const void* packed; int B, C, H, W; cudaStream_t s;
    template <typename T, int MODE> int run() {
        if (packed) {
            const ssm_tensor pd = packed_image_desc(packed, H, W);
            warp_fwd_kernel<T, MODE, true><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(
                cview<T>(&pd), cview<T>(flow), mview<T>(out), 3, make_geom(H, W));
        } else {
            warp_fwd_kernel<T, MODE, false><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(
                cview<T>(img), cview<T>(flow), mview<T>(out), C, make_geom(H, W));
        }
        SSM_LAUNCH_CHECK("ssm_warp_fwd");
        return SSM_OK;
    }
};

struct WarpBwd {
    const ssm_tensor *gout, *img, *flow, *gimg, *gflow; const void* packed; int B, C, H, W; void* ws; cudaStream_t s;
    template <typename T, int MODE> int run() {
        const Geom g = make_geom(H, W);
        const bool want_img = gimg && gimg->data, want_flow = gflow && gflow->data;
        ScatterHdr* hdr = want_img ? (ScatterHdr*)ws : nullptr;
        const long long npx = (long long)H * W;
        long long* acc = want_img ? (long long*)((char*)ws + HDR_BYTES) : nullptr;
        if (want_img) {
            if ((long long)B * C * finalize_chunks(npx) > 2147483647ll)      // before anything is launched
                return fail(SSM_ERR_SHAPE, "ssm_warp_bwd: B*C*H*W too large for one finalise launch");
            cudaError_t e = cudaMemsetAsync(ws, 0, HDR_BYTES + sizeof(long long) * B * C * npx, s);
            if (e != cudaSuccess) return cuda_fail(e, "ssm_warp_bwd memset");
        }
        if (want_flow) {
            if (packed) {
                const ssm_tensor pd = packed_image_desc(packed, H, W);
                warp_bwd_flow_kernel<T, MODE, true><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(
                    cview<T>(gout), cview<T>(&pd), cview<T>(flow), mview<T>(gflow), 3, g, hdr);
            } else {
                warp_bwd_flow_kernel<T, MODE, false><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(
                    cview<T>(gout), cview<T>(img), cview<T>(flow), mview<T>(gflow), C, g, hdr);
            }
            SSM_LAUNCH_CHECK("ssm_warp_bwd (flow)");
        } else if (want_img) {
            absmax_kernel<T><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(cview<T>(gout), C, g, hdr);
            SSM_LAUNCH_CHECK("ssm_warp_bwd (absmax)");
        }
        if (want_img) {
            const int cb = count_bits_for(npx);
            if (C == 3)
                warp_scatter_win_kernel<T, MODE><<<sw_grid(B, H, W), SW_THREADS, 0, s>>>(cview<T>(gout), cview<T>(flow), acc, g, hdr);
            else
                warp_scatter_kernel<T, MODE><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(
                    cview<T>(gout), cview<T>(flow), acc, C, g, hdr, cb);
            SSM_LAUNCH_CHECK("ssm_warp_bwd (scatter)");
            scatter_finalize_kernel<T><<<(unsigned)(B * C) * finalize_chunks(npx), FIN_THREADS, 0, s>>>(acc, nullptr, mview<T>(gimg), C, npx, finalize_chunks(npx), hdr, cb);
            SSM_LAUNCH_CHECK("ssm_warp_bwd (finalize)");
        }
        return SSM_OK;
    }
};

struct PackImage {
    const ssm_tensor* img3; void* packed; int B, H, W; cudaStream_t s;
    template <typename T, int MODE> int run() {
        pack_image_kernel<T><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(cview<T>(img3), (T*)packed, make_geom(H, W));
        SSM_LAUNCH_CHECK("ssm_pack_image");
        return SSM_OK;
    }
};

struct PackFrames {
    const ssm_tensor* img6; void* packed; int B, H, W; cudaStream_t s;
    template <typename T, int MODE> int run() {
        pack_frames_kernel<T><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(cview<T>(img6), (T*)packed, make_geom(H, W));
        SSM_LAUNCH_CHECK("ssm_pack_frames");
        return SSM_OK;
    }
};

struct PackFwd {
    const ssm_tensor* img6; const void* packed; const ssm_tensor* flow4; const float* t; const ssm_tensor* out16;
    int B, N, H, W; cudaStream_t s;
    template <typename T, int MODE, bool PACKED> int run() {
        flow_pack_fwd_kernel<T, MODE, PACKED><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(
            cview<T>(img6), (const T*)packed, cview<T>(flow4), t, mview<T>(out16), N, make_geom(H, W));
        SSM_LAUNCH_CHECK("ssm_flow_pack_fwd");
        return SSM_OK;
    }
};

struct PackFwdNhwc {
    const ssm_tensor* img6; const void* packed; const ssm_tensor* flow4; const float* t; void* out; int out_dtype;
    int B, N, H, W; cudaStream_t s;
    template <typename T, int MODE, bool PACKED> int run() {
        if (out_dtype == SSM_DTYPE_BF16) return go<T, MODE, PACKED, __nv_bfloat16>();
        if constexpr (sizeof(T) == 4) return go<T, MODE, PACKED, float>();
        return fail(SSM_ERR_UNSUPPORTED, "ssm_flow_pack_fwd_nhwc: bf16 inputs with fp32 output is not built");
    }
    template <typename T, int MODE, bool PACKED, typename TO> int go() {
        const long long npx = (long long)H * W;
        View<TO> o;
        o.p = (TO*)out; o.sb = (long long)N * 16 * npx; o.sn = 16 * npx; o.sc = 1;
        flow_pack_fwd_kernel<T, MODE, PACKED, TO, true><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(
            cview<T>(img6), (const T*)packed, cview<T>(flow4), t, o, N, make_geom(H, W));
        SSM_LAUNCH_CHECK("ssm_flow_pack_fwd_nhwc");
        return SSM_OK;
    }
};

struct PackBwd {
    const ssm_tensor *g16, *img6; const void* packed; const ssm_tensor* flow4; const float* t;
    const ssm_tensor *gflow4, *gimg6; int B, N, H, W; void* ws; cudaStream_t s;
    template <typename T, int MODE, bool PACKED> int run() {
        const Geom g = make_geom(H, W);
        const bool want_img = gimg6 && gimg6->data;
        const long long npx = (long long)H * W;
        if (!want_img) {
            flow_pack_bwd_kernel<T, MODE, PACKED, false><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(
                cview<T>(g16), cview<T>(img6), (const T*)packed, cview<T>(flow4), t, mview<T>(gflow4), nullptr, nullptr, N, g);
            SSM_LAUNCH_CHECK("ssm_flow_pack_bwd");
            return SSM_OK;
        }
        ScatterHdr* hdr = (ScatterHdr*)ws;
        long long* acc = (long long*)((char*)ws + HDR_BYTES);
        float* direct = (float*)((char*)ws + HDR_BYTES + sizeof(long long) * B * 6 * npx);
        cudaError_t e = cudaMemsetAsync(ws, 0, HDR_BYTES + sizeof(long long) * B * 6 * npx, s);
        if (e != cudaSuccess) return cuda_fail(e, "ssm_flow_pack_bwd memset");
        flow_pack_bwd_kernel<T, MODE, PACKED, true><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(
            cview<T>(g16), cview<T>(img6), (const T*)packed, cview<T>(flow4), t, mview<T>(gflow4), direct, hdr, N, g);
        SSM_LAUNCH_CHECK("ssm_flow_pack_bwd");
        const int cb = count_bits_for((long long)N * npx);
        flow_pack_scatter_kernel<T, MODE><<<sw_grid(B, H, W), SW_THREADS, 0, s>>>(
            cview<T>(g16), cview<T>(flow4), t, acc, N, g, hdr, cb);
        SSM_LAUNCH_CHECK("ssm_flow_pack_bwd (scatter)");
        scatter_finalize_kernel<T><<<(unsigned)(B * 6) * finalize_chunks(npx), FIN_THREADS, 0, s>>>(acc, direct, mview<T>(gimg6), 6, npx, finalize_chunks(npx), hdr, cb);
        SSM_LAUNCH_CHECK("ssm_flow_pack_bwd (finalize)");
        return SSM_OK;
    }
};

struct FuseFwd {
    const ssm_tensor* img6; const void* packed; const ssm_tensor *flows4, *out5; const float* t;
    const ssm_tensor* out3; int B, N, H, W; cudaStream_t s; bool recomp;
    template <typename T, int MODE, bool PACKED> int run() {
        return recomp ? go<T, MODE, PACKED, true>() : go<T, MODE, PACKED, false>();
    }
    template <typename T, int MODE, bool PACKED, bool RECOMP> int go() {
        fuse_fwd_kernel<T, MODE, PACKED, RECOMP><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(
            cview<T>(img6), (const T*)packed, cview<T>(flows4), cview<T>(out5), t, mview<T>(out3), N, make_geom(H, W));
        SSM_LAUNCH_CHECK("ssm_fuse_fwd");
        return SSM_OK;
    }
};

// fp32 frames and flows with the U-Net output stored in bf16 (ssm_fuse_flow_fwd_mixed)
struct FuseFlowFwdMixed {
    const ssm_tensor* img6; const void* packed; const ssm_tensor *flow4, *out5; const float* t;
    const ssm_tensor* out3; int B, N, H, W; cudaStream_t s;
    template <typename T, int MODE, bool PACKED> int run() {
        fuse_fwd_kernel<float, MODE, PACKED, true, __nv_bfloat16><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(
            cview<float>(img6), (const float*)packed, cview<float>(flow4), cview<__nv_bfloat16>(out5), t,
            mview<float>(out3), N, make_geom(H, W));
        SSM_LAUNCH_CHECK("ssm_fuse_flow_fwd_mixed");
        return SSM_OK;
    }
};

struct FuseBwd {
    const ssm_tensor *g3, *img6; const void* packed; const ssm_tensor *flows4, *out5; const float* t;
    const ssm_tensor *gout5, *gflows4, *gimg6; int B, N, H, W; void* ws; cudaStream_t s; bool recomp;
    template <typename T, int MODE, bool PACKED> int run() {
        return recomp ? go<T, MODE, PACKED, true>() : go<T, MODE, PACKED, false>();
    }
    template <typename T, int MODE, bool PACKED, bool RECOMP> int go() {
        const Geom g = make_geom(H, W);
        const bool want_img = gimg6 && gimg6->data;
        const long long npx = (long long)H * W;
        if (!want_img) {
            fuse_bwd_kernel<T, MODE, PACKED, false, RECOMP><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(
                cview<T>(g3), cview<T>(img6), (const T*)packed, cview<T>(flows4), cview<T>(out5), t, mview<T>(gout5),
                mview<T>(gflows4), nullptr, nullptr, N, g);
            SSM_LAUNCH_CHECK("ssm_fuse_bwd");
            return SSM_OK;
        }
        ScatterHdr* hdr = (ScatterHdr*)ws;
        long long* acc = (long long*)((char*)ws + HDR_BYTES);
        cudaError_t e = cudaMemsetAsync(ws, 0, HDR_BYTES + sizeof(long long) * B * 6 * npx, s);
        if (e != cudaSuccess) return cuda_fail(e, "ssm_fuse_bwd memset");
        fuse_bwd_kernel<T, MODE, PACKED, true, RECOMP><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(
            cview<T>(g3), cview<T>(img6), (const T*)packed, cview<T>(flows4), cview<T>(out5), t, mview<T>(gout5),
            mview<T>(gflows4), nullptr, hdr, N, g);
        SSM_LAUNCH_CHECK("ssm_fuse_bwd");
        const int cb = count_bits_for((long long)N * npx);
        fuse_scatter_kernel<T, MODE, RECOMP><<<sw_grid(B, H, W), SW_THREADS, 0, s>>>(
            cview<T>(g3), cview<T>(flows4), cview<T>(out5), t, acc, N, g, hdr, cb);
        SSM_LAUNCH_CHECK("ssm_fuse_bwd (scatter)");
        scatter_finalize_kernel<T><<<(unsigned)(B * 6) * finalize_chunks(npx), FIN_THREADS, 0, s>>>(acc, nullptr, mview<T>(gimg6), 6, npx, finalize_chunks(npx), hdr, cb);
        SSM_LAUNCH_CHECK("ssm_fuse_bwd (finalize)");
        return SSM_OK;
    }
};

struct FuseLossFwd {
    const ssm_tensor* img6; const void* packed; const ssm_tensor *flow4, *out5, *target; const float* t;
    const ssm_tensor* out3; float* sums; int B, N, H, W, s1, s2; float* partials; cudaStream_t s;
    template <typename T, int MODE, bool PACKED> int run() {
        const unsigned grid = tile_grid(B, H, W);
        fuse_loss_fwd_kernel<T, MODE, PACKED><<<grid, TILE_THREADS, 0, s>>>(
            cview<T>(img6), (const T*)packed, cview<T>(flow4), cview<T>(out5), cview<T>(target), t, mview<T>(out3),
            partials, N, make_geom(H, W), s1, s2);
        SSM_LAUNCH_CHECK("ssm_fuse_loss_fwd");
        loss_reduce_kernel<<<B, 256, 0, s>>>(partials, (int)(grid / B), 2 * N + 1, sums);
        SSM_LAUNCH_CHECK("ssm_fuse_loss_fwd (reduce)");
        return SSM_OK;
    }
};

struct FuseLossBwd {
    const ssm_tensor* g3; const float* gsum; const ssm_tensor* img6; const void* packed;
    const ssm_tensor *flow4, *out5, *target, *out3; const float* t; const ssm_tensor *gout5, *gflow4;
    int B, N, H, W, s1, s2; cudaStream_t s;
    template <typename T, int MODE, bool PACKED> int run() {
        fuse_loss_bwd_kernel<T, MODE, PACKED><<<tile_grid(B, H, W), TILE_THREADS, 0, s>>>(
            cview<T>(g3), gsum, cview<T>(img6), (const T*)packed, cview<T>(flow4), cview<T>(out5), cview<T>(target),
            cview<T>(out3), t, mview<T>(gout5), mview<T>(gflow4), N, make_geom(H, W), s1, s2);
        SSM_LAUNCH_CHECK("ssm_fuse_loss_bwd");
        return SSM_OK;
    }
};

// exhaustive check of div_rn_const against the IEEE division for one divisor.
// out[0] += dividends whose NORMALISED COORDINATE rn(q - 1) differs (what the path consumes,
// layers.py:112); out[1] += dividends whose raw quotient differs; out[2] = one such dividend's bits.
__global__ void selftest_division_kernel(float d, float y, unsigned long long* out) {
    unsigned long long bad_n = 0, bad_q = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < (1ull << 32);
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float s = __uint_as_float((unsigned)i);
        if (!isfinite(s)) continue;
        const float want = __fdiv_rn(s, d);
        const float got = div_rn_const(s, d, y);
        if (__float_as_uint(got) != __float_as_uint(want)) { ++bad_q; out[2] = i; }
        if (__float_as_uint(__fsub_rn(got, 1.0f)) != __float_as_uint(__fsub_rn(want, 1.0f))) ++bad_n;
    }
    if (bad_n) atomicAdd(out, bad_n);
    if (bad_q) atomicAdd(out + 1, bad_q);
}

#define SSM_TRY(expr)            \
    do {                         \
        int rc__ = (expr);       \
        if (rc__ != SSM_OK) return rc__; \
    } while (0)

}  // namespace

extern "C" {

int ssm_version(void) { return SSM_ABI_VERSION; }

const char* ssm_last_error(void) { return g_err; }

size_t ssm_warp_bwd_workspace_bytes(int B, int C, int H, int W) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return HDR_BYTES + sizeof(long long) * (size_t)B * C * H * W;
}
size_t ssm_flow_pack_bwd_workspace_bytes(int B, int N, int H, int W) {
    if (B <= 0 || N <= 0 || H <= 0 || W <= 0) return 0;
    return HDR_BYTES + (sizeof(long long) + sizeof(float)) * (size_t)B * 6 * H * W;
}
size_t ssm_fuse_bwd_workspace_bytes(int B, int N, int H, int W) {
    if (B <= 0 || N <= 0 || H <= 0 || W <= 0) return 0;
    return HDR_BYTES + sizeof(long long) * (size_t)B * 6 * H * W;      // (round 1 added a B x N x 6 fp32 staging buffer)
}

int ssm_selftest_division(int size, unsigned long long* mismatches_device, void* stream) {
    if (size < 1) return fail(SSM_ERR_SHAPE, "size must be >= 1");
    if (!mismatches_device) return fail(SSM_ERR_NULL, "mismatches_device is NULL");
    const float d = (float)(size - 1 > 1 ? size - 1 : 1);
    selftest_division_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(d, 1.0f / d, mismatches_device);
    SSM_LAUNCH_CHECK("ssm_selftest_division");
    return SSM_OK;
}

int ssm_warp_fwd(const ssm_tensor* img, const ssm_tensor* flow, const ssm_tensor* out,
                 int B, int C, int H, int W, int dtype, int coord_mode, void* stream) {
    SSM_TRY(check_common(B, 1, C, H, W, dtype, coord_mode));
    SSM_TRY(check_tensor(img, "img", dtype, true));
    SSM_TRY(check_tensor(flow, "flow", dtype, true));
    SSM_TRY(check_tensor(out, "out", dtype, true));
    return dispatch(dtype, coord_mode, WarpFwd{img, flow, out, nullptr, B, C, H, W, (cudaStream_t)stream});
}

size_t ssm_packed_image_bytes(int B, int H, int W, int dtype) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)B * H * W * 4 * (dtype == SSM_DTYPE_F32 ? 4 : 2);
}

int ssm_pack_image(const ssm_tensor* img3, void* packed, int B, int H, int W, int dtype, void* stream) {
    SSM_TRY(check_common(B, 1, 3, H, W, dtype, SSM_COORD_DIV));
    SSM_TRY(check_tensor(img3, "img3", dtype, true));
    if (!packed) return fail(SSM_ERR_NULL, "packed is NULL");
    if (((uintptr_t)packed) % 16 != 0) return fail(SSM_ERR_ALIGN, "packed must be 16-byte aligned");
    return dispatch(dtype, SSM_COORD_DIV, PackImage{img3, packed, B, H, W, (cudaStream_t)stream});
}

int ssm_warp_fwd_packed(const void* packed, const ssm_tensor* flow, const ssm_tensor* out,
                        int B, int H, int W, int dtype, int coord_mode, void* stream) {
    SSM_TRY(check_common(B, 1, 3, H, W, dtype, coord_mode));
    if (!packed) return fail(SSM_ERR_NULL, "packed is NULL");
    if (((uintptr_t)packed) % 16 != 0) return fail(SSM_ERR_ALIGN, "packed must be 16-byte aligned");
    SSM_TRY(check_tensor(flow, "flow", dtype, true));
    SSM_TRY(check_tensor(out, "out", dtype, true));
    return dispatch(dtype, coord_mode, WarpFwd{nullptr, flow, out, packed, B, 3, H, W, (cudaStream_t)stream});
}

int ssm_warp_bwd(const ssm_tensor* grad_out, const ssm_tensor* img, const ssm_tensor* flow,
                 const ssm_tensor* grad_img, const ssm_tensor* grad_flow,
                 int B, int C, int H, int W, int dtype, int coord_mode,
                 void* workspace, size_t workspace_bytes, void* stream) {
    SSM_TRY(check_common(B, 1, C, H, W, dtype, coord_mode));
    SSM_TRY(check_tensor(grad_out, "grad_out", dtype, true));
    SSM_TRY(check_tensor(img, "img", dtype, true));
    SSM_TRY(check_tensor(flow, "flow", dtype, true));
    SSM_TRY(check_tensor(grad_img, "grad_img", dtype, false));
    SSM_TRY(check_tensor(grad_flow, "grad_flow", dtype, false));
    if (grad_img && grad_img->data) {
        if (!workspace || workspace_bytes < ssm_warp_bwd_workspace_bytes(B, C, H, W))
            return fail(SSM_ERR_WORKSPACE, "ssm_warp_bwd: grad_img needs %zu workspace bytes, got %zu",
                        ssm_warp_bwd_workspace_bytes(B, C, H, W), workspace ? workspace_bytes : (size_t)0);
        if (((uintptr_t)workspace) % 16 != 0) return fail(SSM_ERR_ALIGN, "workspace must be 16-byte aligned");
    }
    return dispatch(dtype, coord_mode,
                    WarpBwd{grad_out, img, flow, grad_img, grad_flow, nullptr, B, C, H, W, workspace, (cudaStream_t)stream});
}

int ssm_warp_bwd_packed(const ssm_tensor* grad_out, const void* packed, const ssm_tensor* flow,
                        const ssm_tensor* grad_img, const ssm_tensor* grad_flow,
                        int B, int H, int W, int dtype, int coord_mode,
                        void* workspace, size_t workspace_bytes, void* stream) {
    SSM_TRY(check_common(B, 1, 3, H, W, dtype, coord_mode));
    if (!packed) return fail(SSM_ERR_NULL, "packed is NULL");
    if (((uintptr_t)packed) % 16 != 0) return fail(SSM_ERR_ALIGN, "packed must be 16-byte aligned");
    SSM_TRY(check_tensor(grad_out, "grad_out", dtype, true));
    SSM_TRY(check_tensor(flow, "flow", dtype, true));
    SSM_TRY(check_tensor(grad_img, "grad_img", dtype, false));
    SSM_TRY(check_tensor(grad_flow, "grad_flow", dtype, false));
    if (grad_img && grad_img->data) {
        if (!workspace || workspace_bytes < ssm_warp_bwd_workspace_bytes(B, 3, H, W))
            return fail(SSM_ERR_WORKSPACE, "ssm_warp_bwd_packed: grad_img needs %zu workspace bytes, got %zu",
                        ssm_warp_bwd_workspace_bytes(B, 3, H, W), workspace ? workspace_bytes : (size_t)0);
        if (((uintptr_t)workspace) % 16 != 0) return fail(SSM_ERR_ALIGN, "workspace must be 16-byte aligned");
    }
    return dispatch(dtype, coord_mode,
                    WarpBwd{grad_out, nullptr, flow, grad_img, grad_flow, packed, B, 3, H, W, workspace, (cudaStream_t)stream});
}

size_t ssm_packed_frames_bytes(int B, int H, int W, int dtype) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)B * 2 * H * W * 4 * (dtype == SSM_DTYPE_F32 ? 4 : 2);
}

int ssm_pack_frames(const ssm_tensor* img6, void* packed, int B, int H, int W, int dtype, void* stream) {
    SSM_TRY(check_common(B, 1, 6, H, W, dtype, SSM_COORD_DIV));
    SSM_TRY(check_tensor(img6, "img6", dtype, true));
    if (!packed) return fail(SSM_ERR_NULL, "packed is NULL");
    if (((uintptr_t)packed) % 16 != 0) return fail(SSM_ERR_ALIGN, "packed must be 16-byte aligned");
    return dispatch(dtype, SSM_COORD_DIV, PackFrames{img6, packed, B, H, W, (cudaStream_t)stream});
}

int ssm_flow_pack_fwd(const ssm_tensor* img6, const void* packed, const ssm_tensor* flow4, const float* t,
                      const ssm_tensor* out16, int B, int N, int H, int W,
                      int dtype, int coord_mode, void* stream) {
    SSM_TRY(check_packed(packed));
    SSM_TRY(check_common(B, N, 16, H, W, dtype, coord_mode));
    SSM_TRY(check_tensor(img6, "img6", dtype, true));
    SSM_TRY(check_tensor(flow4, "flow4", dtype, true));
    SSM_TRY(check_tensor(out16, "out16", dtype, true));
    if (!t) return fail(SSM_ERR_NULL, "t is NULL");
    return dispatch3(dtype, coord_mode, packed != nullptr, PackFwd{img6, packed, flow4, t, out16, B, N, H, W, (cudaStream_t)stream});
}

int ssm_flow_pack_fwd_nhwc(const ssm_tensor* img6, const void* packed, const ssm_tensor* flow4, const float* t,
                           void* out16_nhwc, int B, int N, int H, int W,
                           int dtype, int out_dtype, int coord_mode, void* stream) {
    SSM_TRY(check_packed(packed));
    SSM_TRY(check_common(B, N, 16, H, W, dtype, coord_mode));
    if (out_dtype != SSM_DTYPE_F32 && out_dtype != SSM_DTYPE_BF16) return fail(SSM_ERR_DTYPE, "unknown out_dtype %d", out_dtype);
    SSM_TRY(check_tensor(img6, "img6", dtype, true));
    SSM_TRY(check_tensor(flow4, "flow4", dtype, true));
    if (!out16_nhwc) return fail(SSM_ERR_NULL, "out16_nhwc is NULL");
    if (((uintptr_t)out16_nhwc) % 32 != 0) return fail(SSM_ERR_ALIGN, "out16_nhwc must be 32-byte aligned");
    if (!t) return fail(SSM_ERR_NULL, "t is NULL");
    return dispatch3(dtype, coord_mode, packed != nullptr,
                     PackFwdNhwc{img6, packed, flow4, t, out16_nhwc, out_dtype, B, N, H, W, (cudaStream_t)stream});
}

int ssm_flow_pack_bwd(const ssm_tensor* grad16, const ssm_tensor* img6, const void* packed,
                      const ssm_tensor* flow4, const float* t, const ssm_tensor* grad_flow4,
                      const ssm_tensor* grad_img6, int B, int N, int H, int W, int dtype, int coord_mode,
                      void* workspace, size_t workspace_bytes, void* stream) {
    SSM_TRY(check_packed(packed));
    SSM_TRY(check_common(B, N, 16, H, W, dtype, coord_mode));
    SSM_TRY(check_tensor(grad16, "grad16", dtype, true));
    SSM_TRY(check_tensor(img6, "img6", dtype, true));
    SSM_TRY(check_tensor(flow4, "flow4", dtype, true));
    SSM_TRY(check_tensor(grad_flow4, "grad_flow4", dtype, false));
    SSM_TRY(check_tensor(grad_img6, "grad_img6", dtype, false));
    if (!t) return fail(SSM_ERR_NULL, "t is NULL");
    if (grad_img6 && grad_img6->data) {
        if (!workspace || workspace_bytes < ssm_flow_pack_bwd_workspace_bytes(B, N, H, W))
            return fail(SSM_ERR_WORKSPACE, "ssm_flow_pack_bwd: grad_img6 needs %zu workspace bytes, got %zu",
                        ssm_flow_pack_bwd_workspace_bytes(B, N, H, W), workspace ? workspace_bytes : (size_t)0);
        if (((uintptr_t)workspace) % 16 != 0) return fail(SSM_ERR_ALIGN, "workspace must be 16-byte aligned");
    }
    return dispatch3(dtype, coord_mode, packed != nullptr,
                     PackBwd{grad16, img6, packed, flow4, t, grad_flow4, grad_img6, B, N, H, W, workspace, (cudaStream_t)stream});
}

static int fuse_fwd_impl(const ssm_tensor* img6, const void* packed, const ssm_tensor* flows4, const ssm_tensor* out5,
                         const float* t, const ssm_tensor* out3, int B, int N, int H, int W,
                         int dtype, int coord_mode, void* stream, bool recomp) {
    SSM_TRY(check_packed(packed));
    SSM_TRY(check_common(B, N, 5, H, W, dtype, coord_mode));
    SSM_TRY(check_tensor(img6, "img6", dtype, true));
    SSM_TRY(check_tensor(flows4, recomp ? "flow4" : "flows4", dtype, true));
    SSM_TRY(check_tensor(out5, "out5", dtype, true));
    SSM_TRY(check_tensor(out3, "out3", dtype, true));
    if (!t) return fail(SSM_ERR_NULL, "t is NULL");
    return dispatch3(dtype, coord_mode, packed != nullptr,
                     FuseFwd{img6, packed, flows4, out5, t, out3, B, N, H, W, (cudaStream_t)stream, recomp});
}

static int fuse_bwd_impl(const ssm_tensor* grad3, const ssm_tensor* img6, const void* packed,
                         const ssm_tensor* flows4, const ssm_tensor* out5, const float* t, const ssm_tensor* grad_out5,
                         const ssm_tensor* grad_flows4, const ssm_tensor* grad_img6,
                         int B, int N, int H, int W, int dtype, int coord_mode,
                         void* workspace, size_t workspace_bytes, void* stream, bool recomp) {
    SSM_TRY(check_packed(packed));
    SSM_TRY(check_common(B, N, 5, H, W, dtype, coord_mode));
    SSM_TRY(check_tensor(grad3, "grad3", dtype, true));
    SSM_TRY(check_tensor(img6, "img6", dtype, true));
    SSM_TRY(check_tensor(flows4, recomp ? "flow4" : "flows4", dtype, true));
    SSM_TRY(check_tensor(out5, "out5", dtype, true));
    SSM_TRY(check_tensor(grad_out5, "grad_out5", dtype, false));
    SSM_TRY(check_tensor(grad_flows4, recomp ? "grad_flow4" : "grad_flows4", dtype, false));
    SSM_TRY(check_tensor(grad_img6, "grad_img6", dtype, false));
    if (!t) return fail(SSM_ERR_NULL, "t is NULL");
    if (grad_img6 && grad_img6->data) {
        if (!workspace || workspace_bytes < ssm_fuse_bwd_workspace_bytes(B, N, H, W))
            return fail(SSM_ERR_WORKSPACE, "ssm_fuse_bwd: grad_img6 needs %zu workspace bytes, got %zu",
                        ssm_fuse_bwd_workspace_bytes(B, N, H, W), workspace ? workspace_bytes : (size_t)0);
        if (((uintptr_t)workspace) % 16 != 0) return fail(SSM_ERR_ALIGN, "workspace must be 16-byte aligned");
    }
    return dispatch3(dtype, coord_mode, packed != nullptr,
                     FuseBwd{grad3, img6, packed, flows4, out5, t, grad_out5, grad_flows4, grad_img6, B, N, H, W,
                             workspace, (cudaStream_t)stream, recomp});
}

int ssm_fuse_fwd(const ssm_tensor* img6, const void* packed, const ssm_tensor* flows4, const ssm_tensor* out5,
                 const float* t, const ssm_tensor* out3, int B, int N, int H, int W,
                 int dtype, int coord_mode, void* stream) {
    return fuse_fwd_impl(img6, packed, flows4, out5, t, out3, B, N, H, W, dtype, coord_mode, stream, false);
}

int ssm_fuse_bwd(const ssm_tensor* grad3, const ssm_tensor* img6, const void* packed,
                 const ssm_tensor* flows4, const ssm_tensor* out5, const float* t, const ssm_tensor* grad_out5,
                 const ssm_tensor* grad_flows4, const ssm_tensor* grad_img6,
                 int B, int N, int H, int W, int dtype, int coord_mode,
                 void* workspace, size_t workspace_bytes, void* stream) {
    return fuse_bwd_impl(grad3, img6, packed, flows4, out5, t, grad_out5, grad_flows4, grad_img6, B, N, H, W, dtype,
                         coord_mode, workspace, workspace_bytes, stream, false);
}

int ssm_fuse_flow_fwd(const ssm_tensor* img6, const void* packed, const ssm_tensor* flow4, const ssm_tensor* out5,
                      const float* t, const ssm_tensor* out3, int B, int N, int H, int W,
                      int dtype, int coord_mode, void* stream) {
    return fuse_fwd_impl(img6, packed, flow4, out5, t, out3, B, N, H, W, dtype, coord_mode, stream, true);
}

int ssm_fuse_flow_fwd_mixed(const ssm_tensor* img6, const void* packed, const ssm_tensor* flow4,
                            const ssm_tensor* out5, int out5_dtype, const float* t, const ssm_tensor* out3,
                            int B, int N, int H, int W, int dtype, int coord_mode, void* stream) {
    if (out5_dtype == dtype)
        return fuse_fwd_impl(img6, packed, flow4, out5, t, out3, B, N, H, W, dtype, coord_mode, stream, true);
    if (dtype != SSM_DTYPE_F32 || out5_dtype != SSM_DTYPE_BF16)
        return fail(SSM_ERR_UNSUPPORTED, "ssm_fuse_flow_fwd_mixed: only fp32 frames/flows with a bf16 out5 (got dtype %d, out5_dtype %d)",
                    dtype, out5_dtype);
    SSM_TRY(check_packed(packed));
    SSM_TRY(check_common(B, N, 5, H, W, dtype, coord_mode));
    SSM_TRY(check_tensor(img6, "img6", dtype, true));
    SSM_TRY(check_tensor(flow4, "flow4", dtype, true));
    SSM_TRY(check_tensor(out5, "out5", out5_dtype, true));
    SSM_TRY(check_tensor(out3, "out3", dtype, true));
    if (!t) return fail(SSM_ERR_NULL, "t is NULL");
    return dispatch3(dtype, coord_mode, packed != nullptr,
                     FuseFlowFwdMixed{img6, packed, flow4, out5, t, out3, B, N, H, W, (cudaStream_t)stream});
}

int ssm_fuse_flow_bwd(const ssm_tensor* grad3, const ssm_tensor* img6, const void* packed,
                      const ssm_tensor* flow4, const ssm_tensor* out5, const float* t, const ssm_tensor* grad_out5,
                      const ssm_tensor* grad_flow4, const ssm_tensor* grad_img6,
                      int B, int N, int H, int W, int dtype, int coord_mode,
                      void* workspace, size_t workspace_bytes, void* stream) {
    return fuse_bwd_impl(grad3, img6, packed, flow4, out5, t, grad_out5, grad_flow4, grad_img6, B, N, H, W, dtype,
                         coord_mode, workspace, workspace_bytes, stream, true);
}

size_t ssm_fuse_loss_workspace_bytes(int B, int N, int H, int W) {
    if (B <= 0 || N <= 0 || H <= 0 || W <= 0) return 0;
    return sizeof(float) * (size_t)tile_grid(B, H, W) * (2 * (size_t)N + 1);
}

int ssm_fuse_loss_fwd(const ssm_tensor* img6, const void* packed, const ssm_tensor* flow4, const ssm_tensor* out5,
                      const ssm_tensor* target, const float* t, const ssm_tensor* out3, float* sums,
                      int B, int N, int H, int W, int dtype, int coord_mode, int stage1_loss, int stage2_loss,
                      void* workspace, size_t workspace_bytes, void* stream) {
    SSM_TRY(check_packed(packed));
    SSM_TRY(check_common(B, N, 5, H, W, dtype, coord_mode));
    SSM_TRY(check_tensor(img6, "img6", dtype, true));
    SSM_TRY(check_tensor(flow4, "flow4", dtype, true));
    SSM_TRY(check_tensor(out5, "out5", dtype, true));
    SSM_TRY(check_tensor(target, "target", dtype, true));
    SSM_TRY(check_tensor(out3, "out3", dtype, true));
    if (!t) return fail(SSM_ERR_NULL, "t is NULL");
    if (!sums) return fail(SSM_ERR_NULL, "sums is NULL");
    if (!workspace || workspace_bytes < ssm_fuse_loss_workspace_bytes(B, N, H, W))
        return fail(SSM_ERR_WORKSPACE, "ssm_fuse_loss_fwd: needs %zu workspace bytes, got %zu",
                    ssm_fuse_loss_workspace_bytes(B, N, H, W), workspace ? workspace_bytes : (size_t)0);
    if (((uintptr_t)workspace) % 4 != 0 || ((uintptr_t)sums) % 4 != 0) return fail(SSM_ERR_ALIGN, "workspace / sums must be 4-byte aligned");
    return dispatch3(dtype, coord_mode, packed != nullptr,
                     FuseLossFwd{img6, packed, flow4, out5, target, t, out3, sums, B, N, H, W, stage1_loss != 0,
                                 stage2_loss != 0, (float*)workspace, (cudaStream_t)stream});
}

int ssm_fuse_loss_bwd(const ssm_tensor* grad3, const float* grad_sums, const ssm_tensor* img6, const void* packed,
                      const ssm_tensor* flow4, const ssm_tensor* out5, const ssm_tensor* target, const ssm_tensor* out3,
                      const float* t, const ssm_tensor* grad_out5, const ssm_tensor* grad_flow4,
                      int B, int N, int H, int W, int dtype, int coord_mode, int stage1_loss, int stage2_loss,
                      void* stream) {
    SSM_TRY(check_packed(packed));
    SSM_TRY(check_common(B, N, 5, H, W, dtype, coord_mode));
    SSM_TRY(check_tensor(grad3, "grad3", dtype, false));
    SSM_TRY(check_tensor(img6, "img6", dtype, true));
    SSM_TRY(check_tensor(flow4, "flow4", dtype, true));
    SSM_TRY(check_tensor(out5, "out5", dtype, true));
    SSM_TRY(check_tensor(target, "target", dtype, true));
    SSM_TRY(check_tensor(out3, "out3", dtype, true));
    SSM_TRY(check_tensor(grad_out5, "grad_out5", dtype, false));
    SSM_TRY(check_tensor(grad_flow4, "grad_flow4", dtype, false));
    if (!t) return fail(SSM_ERR_NULL, "t is NULL");
    if (!grad_sums) return fail(SSM_ERR_NULL, "grad_sums is NULL");
    return dispatch3(dtype, coord_mode, packed != nullptr,
                     FuseLossBwd{grad3, grad_sums, img6, packed, flow4, out5, target, out3, t, grad_out5, grad_flow4,
                                 B, N, H, W, stage1_loss != 0, stage2_loss != 0, (cudaStream_t)stream});
}

int ssm_frames_from_u8(const unsigned char* src, long long src_frame_stride, int src_row_stride, int bgr,
                       int F, int H_in, int W_in, int H, int W, int top, int left,
                       const float* lut_device, const float* pad_value3, const ssm_tensor* planar, void* rgbx,
                       int dtype, void* stream) {
    if (F <= 0 || H_in <= 0 || W_in <= 0 || H <= 0 || W <= 0)
        return fail(SSM_ERR_SHAPE, "ssm_frames_from_u8: F, H_in, W_in, H, W must be positive");
    if (top < 0 || left < 0 || top + H_in > H || left + W_in > W)
        return fail(SSM_ERR_SHAPE, "ssm_frames_from_u8: the %d x %d source at (%d, %d) does not fit %d x %d", H_in, W_in, top, left, H, W);
    if (W % 4 != 0) return fail(SSM_ERR_SHAPE, "ssm_frames_from_u8: padded width must be a multiple of 4 (got %d)", W);
    if ((long long)H * W > MAX_PLANE) return fail(SSM_ERR_SHAPE, "H*W too large (%d x %d)", H, W);
    if (dtype != SSM_DTYPE_F32 && dtype != SSM_DTYPE_BF16) return fail(SSM_ERR_DTYPE, "unknown dtype %d", dtype);
    if (!src || !lut_device || !pad_value3) return fail(SSM_ERR_NULL, "ssm_frames_from_u8: src, lut_device or pad_value3 is NULL");
    const bool want_planar = planar && planar->data;
    if (!want_planar && !rgbx) return fail(SSM_ERR_NULL, "ssm_frames_from_u8: neither planar nor rgbx requested");
    SSM_TRY(check_tensor(planar, "planar", dtype, false));
    if (want_planar && ((uintptr_t)planar->data % 16 != 0 || planar->stride_b % 4 != 0 || planar->stride_c % 4 != 0) && dtype == SSM_DTYPE_F32)
        return fail(SSM_ERR_ALIGN, "ssm_frames_from_u8: planar must be 16-byte aligned with strides multiple of 4");
    SSM_TRY(check_packed(rgbx));
    const long long quads = (long long)F * H * (W / 4);
    const unsigned grid = (unsigned)((quads + 255) / 256 < 148ll * 32 ? (quads + 255) / 256 : 148ll * 32);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == SSM_DTYPE_F32)
        frames_from_u8_kernel<float><<<grid, 256, 0, s>>>(src, src_frame_stride, src_row_stride, bgr != 0, H_in, W_in, H, W, top, left,
            lut_device, pad_value3[0], pad_value3[1], pad_value3[2], mview<float>(planar), (float*)rgbx, quads);
    else
        frames_from_u8_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(src, src_frame_stride, src_row_stride, bgr != 0, H_in, W_in, H, W, top, left,
            lut_device, pad_value3[0], pad_value3[1], pad_value3[2], mview<__nv_bfloat16>(planar), (__nv_bfloat16*)rgbx, quads);
    SSM_LAUNCH_CHECK("ssm_frames_from_u8");
    return SSM_OK;
}

int ssm_frames_to_u8(const ssm_tensor* planar, int F, int H, int W, int top, int left, int H_out, int W_out,
                     const float* mean3, const float* std3, float scale, int bgr, int saturate,
                     unsigned char* dst, long long dst_frame_stride, int dst_row_stride, int dtype, void* stream) {
    if (F <= 0 || H <= 0 || W <= 0 || H_out <= 0 || W_out <= 0)
        return fail(SSM_ERR_SHAPE, "ssm_frames_to_u8: F, H, W, H_out, W_out must be positive");
    if (top < 0 || left < 0 || top + H_out > H || left + W_out > W)
        return fail(SSM_ERR_SHAPE, "ssm_frames_to_u8: the %d x %d crop at (%d, %d) leaves %d x %d", H_out, W_out, top, left, H, W);
    if (dtype != SSM_DTYPE_F32 && dtype != SSM_DTYPE_BF16) return fail(SSM_ERR_DTYPE, "unknown dtype %d", dtype);
    if (!dst || !mean3 || !std3) return fail(SSM_ERR_NULL, "ssm_frames_to_u8: dst, mean3 or std3 is NULL");
    SSM_TRY(check_tensor(planar, "planar", dtype, true));
    const long long total = (long long)F * H_out * W_out;
    const unsigned grid = (unsigned)((total + 255) / 256 < 148ll * 32 ? (total + 255) / 256 : 148ll * 32);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == SSM_DTYPE_F32)
        frames_to_u8_kernel<float><<<grid, 256, 0, s>>>(cview<float>(planar), H, W, top, left, H_out, W_out, mean3[0], mean3[1], mean3[2],
            std3[0], std3[1], std3[2], scale, bgr != 0, saturate != 0, dst, dst_frame_stride, dst_row_stride, total);
    else
        frames_to_u8_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(cview<__nv_bfloat16>(planar), H, W, top, left, H_out, W_out, mean3[0], mean3[1], mean3[2],
            std3[0], std3[1], std3[2], scale, bgr != 0, saturate != 0, dst, dst_frame_stride, dst_row_stride, total);
    SSM_LAUNCH_CHECK("ssm_frames_to_u8");
    return SSM_OK;
}

// ---------------------------------------------------------------------------------------------
// Element-wise steps between the U-Nets' cuDNN convolutions (channels-last activations, inference)
static int check_glue(const char* who, const void* a, const void* b, long long count, int C, int dtype) {
    if (!a || !b) return fail(SSM_ERR_NULL, "%s: NULL pointer", who);
    if (count <= 0 || C <= 0) return fail(SSM_ERR_SHAPE, "%s: sizes must be positive", who);
    if (C % 8 != 0) return fail(SSM_ERR_SHAPE, "%s: C must be a multiple of 8 (got %d)", who, C);
    if (dtype != SSM_DTYPE_F32 && dtype != SSM_DTYPE_BF16) return fail(SSM_ERR_DTYPE, "unknown dtype %d", dtype);
    if (((uintptr_t)a | (uintptr_t)b) % 16 != 0) return fail(SSM_ERR_ALIGN, "%s: pointers must be 16-byte aligned", who);
    if ((count * (C / 8) + 255) / 256 > 2147483647ll) return fail(SSM_ERR_SHAPE, "%s: too large for one launch", who);
    return SSM_OK;
}

int ssm_upsample2x_nhwc(const void* in, void* out, int M, int H, int W, int C, long long out_pixel_stride, int dtype, void* stream) {
    if (M <= 0 || H <= 0 || W <= 0) return fail(SSM_ERR_SHAPE, "ssm_upsample2x_nhwc: M, H, W must be positive");
    SSM_TRY(check_glue("ssm_upsample2x_nhwc", in, out, (long long)M * H * W, C, dtype));
    if (out_pixel_stride < C || out_pixel_stride % 8 != 0)
        return fail(SSM_ERR_SHAPE, "ssm_upsample2x_nhwc: out_pixel_stride must be a multiple of 8 and >= C");
    const long long total = (long long)M * ((H + UPS_ROWS - 1) / UPS_ROWS) * W * (C / 8);     // one thread per row block
    const unsigned grid = (unsigned)((total + UPS_BLOCK - 1) / UPS_BLOCK);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == SSM_DTYPE_F32)
        upsample2x_nhwc_kernel<float><<<grid, UPS_BLOCK, 0, s>>>((const float*)in, (float*)out, H, W, C / 8, out_pixel_stride, total);
    else
        upsample2x_nhwc_kernel<__nv_bfloat16><<<grid, UPS_BLOCK, 0, s>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, H, W, C / 8, out_pixel_stride, total);
    SSM_LAUNCH_CHECK("ssm_upsample2x_nhwc");
    return SSM_OK;
}

int ssm_bias_leaky_nhwc_to(const void* y, const float* bias, long long pixels, int C, float slope,
                           void* out1, long long out1_pixel_stride, void* out2, long long out2_pixel_stride,
                           int dtype, void* stream) {
    SSM_TRY(check_glue("ssm_bias_leaky_nhwc", y, bias, pixels, C, dtype));
    SSM_TRY(check_glue("ssm_bias_leaky_nhwc", out1, out2 ? out2 : out1, pixels, C, dtype));
    if (out1_pixel_stride < C || out1_pixel_stride % 8 != 0 || (out2 && (out2_pixel_stride < C || out2_pixel_stride % 8 != 0)))
        return fail(SSM_ERR_SHAPE, "ssm_bias_leaky_nhwc: output pixel strides must be multiples of 8 and >= C");
    const long long total = pixels * (C / 8);
    const unsigned grid = (unsigned)((total + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == SSM_DTYPE_F32)
        bias_leaky_nhwc_kernel<float><<<grid, 256, 0, s>>>((const float*)y, bias, C / 8, slope, total, (float*)out1, out1_pixel_stride,
                                                           (float*)out2, out2_pixel_stride);
    else
        bias_leaky_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)y, bias, C / 8, slope, total, (__nv_bfloat16*)out1,
                                                                   out1_pixel_stride, (__nv_bfloat16*)out2, out2_pixel_stride);
    SSM_LAUNCH_CHECK("ssm_bias_leaky_nhwc");
    return SSM_OK;
}

int ssm_bias_leaky_nhwc(void* y, const float* bias, long long pixels, int C, float slope, int dtype, void* stream) {
    return ssm_bias_leaky_nhwc_to(y, bias, pixels, C, slope, y, C, nullptr, 0, dtype, stream);
}

int ssm_avgpool2_nhwc(const void* in, void* out, int M, int H_out, int W_out, int C, int dtype, void* stream) {
    if (M <= 0 || H_out <= 0 || W_out <= 0) return fail(SSM_ERR_SHAPE, "ssm_avgpool2_nhwc: M, H_out, W_out must be positive");
    SSM_TRY(check_glue("ssm_avgpool2_nhwc", in, out, (long long)M * H_out * W_out, C, dtype));
    const long long total = (long long)M * H_out * W_out * (C / 8);
    const unsigned grid = (unsigned)((total + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == SSM_DTYPE_F32)
        avgpool2_nhwc_kernel<float><<<grid, 256, 0, s>>>((const float*)in, (float*)out, H_out, W_out, C / 8, total);
    else
        avgpool2_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, H_out, W_out, C / 8, total);
    SSM_LAUNCH_CHECK("ssm_avgpool2_nhwc");
    return SSM_OK;
}

int ssm_upsample2x_bwd_nhwc(const void* grad_out, void* grad_in, int M, int H, int W, int C, long long grad_out_pixel_stride,
                            int dtype, void* stream) {
    if (M <= 0 || H <= 0 || W <= 0) return fail(SSM_ERR_SHAPE, "ssm_upsample2x_bwd_nhwc: M, H, W must be positive");
    SSM_TRY(check_glue("ssm_upsample2x_bwd_nhwc", grad_out, grad_in, (long long)M * H * W, C, dtype));
    if (grad_out_pixel_stride < C || grad_out_pixel_stride % 8 != 0)
        return fail(SSM_ERR_SHAPE, "ssm_upsample2x_bwd_nhwc: grad_out_pixel_stride must be a multiple of 8 and >= C");
    const long long total = (long long)M * H * W * (C / 8);
    const unsigned grid = (unsigned)((total + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == SSM_DTYPE_F32)
        upsample2x_bwd_nhwc_kernel<float><<<grid, 256, 0, s>>>((const float*)grad_out, (float*)grad_in, H, W, C / 8, grad_out_pixel_stride, total);
    else
        upsample2x_bwd_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)grad_out, (__nv_bfloat16*)grad_in, H, W, C / 8,
                                                                      grad_out_pixel_stride, total);
    SSM_LAUNCH_CHECK("ssm_upsample2x_bwd_nhwc");
    return SSM_OK;
}

int ssm_leaky_bwd_nhwc(const void* grad_y, const void* y, void* grad_x, long long pixels, int C, float slope, int dtype, void* stream) {
    SSM_TRY(check_glue("ssm_leaky_bwd_nhwc", grad_y, y, pixels, C, dtype));
    SSM_TRY(check_glue("ssm_leaky_bwd_nhwc", grad_x, y, pixels, C, dtype));
    const long long total = pixels * (C / 8);
    const unsigned grid = (unsigned)((total + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == SSM_DTYPE_F32)
        leaky_bwd_nhwc_kernel<float><<<grid, 256, 0, s>>>((const float*)grad_y, (const float*)y, (float*)grad_x, slope, total);
    else
        leaky_bwd_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)grad_y, (const __nv_bfloat16*)y, (__nv_bfloat16*)grad_x, slope, total);
    SSM_LAUNCH_CHECK("ssm_leaky_bwd_nhwc");
    return SSM_OK;
}

int ssm_avgpool2_bwd_nhwc(const void* grad_out, void* grad_in, int M, int H_out, int W_out, int C, int dtype, void* stream) {
    if (M <= 0 || H_out <= 0 || W_out <= 0) return fail(SSM_ERR_SHAPE, "ssm_avgpool2_bwd_nhwc: M, H_out, W_out must be positive");
    SSM_TRY(check_glue("ssm_avgpool2_bwd_nhwc", grad_out, grad_in, (long long)M * H_out * W_out, C, dtype));
    const long long total = (long long)M * H_out * W_out * (C / 8);
    const unsigned grid = (unsigned)((total + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == SSM_DTYPE_F32)
        avgpool2_bwd_nhwc_kernel<float><<<grid, 256, 0, s>>>((const float*)grad_out, (float*)grad_in, H_out, W_out, C / 8, total);
    else
        avgpool2_bwd_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)grad_out, (__nv_bfloat16*)grad_in, H_out, W_out, C / 8, total);
    SSM_LAUNCH_CHECK("ssm_avgpool2_bwd_nhwc");
    return SSM_OK;
}

// ---------------------------------------------------------------------------------------------
// 8-bit frames: entry tables + the forward kernels that gather from them (ssm_q8.cuh)
static int check_q8(const char* who, int B, int N, int H, int W, int coord_mode, const void* quads, const float* norm6) {
    SSM_TRY(check_common(B, N, 16, H, W, SSM_DTYPE_F32, coord_mode));
    if (W % 2 != 0) return fail(SSM_ERR_SHAPE, "%s: W must be even (got %d): two pixels per thread", who, W);
    if (!quads || !norm6) return fail(SSM_ERR_NULL, "%s: quads or norm6 is NULL", who);
    if (((uintptr_t)quads) % 16 != 0) return fail(SSM_ERR_ALIGN, "%s: quads must be 16-byte aligned", who);
    const long long tiles = (long long)B * ((H + Q8_TILE_H - 1) / Q8_TILE_H) * ((W + Q8_TILE_W - 1) / Q8_TILE_W);
    if (tiles > 2147483647ll) return fail(SSM_ERR_SHAPE, "%s: too many tiles for one launch", who);
    if ((long long)(H + 1) * (W + 1) > MAX_PLANE) return fail(SSM_ERR_SHAPE, "%s: H*W too large", who);
    // the gathers address the launch's entry tables with an unsigned 32-bit entry index
    if ((long long)B * 2 * (H + 1) * (W + 1) > 4294967295ll)
        return fail(SSM_ERR_SHAPE, "%s: B*2*(H+1)*(W+1) = %lld entries exceed 2^32 - 1: split the batch", who, (long long)B * 2 * (H + 1) * (W + 1));
    // floor() by the 1.5 * 2^23 addition needs every in-frame coordinate below 2^22
    if (H >= (1 << 22) || W >= (1 << 22)) return fail(SSM_ERR_SHAPE, "%s: H and W must be below 2^22", who);
    return SSM_OK;
}
// 8-byte vector access: base and every stride a multiple of 2 elements (fp32) / the pair of bf16 values aligned to 4 bytes
static int check_pairable(const ssm_tensor* t, const char* name, int dtype) {
    const size_t esz = dtype == SSM_DTYPE_F32 ? 4 : 2;
    if (((uintptr_t)t->data) % (2 * esz) != 0 || t->stride_b % 2 != 0 || t->stride_n % 2 != 0 || t->stride_c % 2 != 0)
        return fail(SSM_ERR_ALIGN, "%s: base must be aligned to two elements and strides must be even (two pixels per access)", name);
    return SSM_OK;
}
static Norm3 make_norm(const float* norm6) {
    Norm3 n;
    for (int k = 0; k < 3; ++k) { n.a[k] = norm6[k]; n.c[k] = norm6[3 + k]; }
    return n;
}
static unsigned q8_grid(int B, int H, int W) {
    return (unsigned)((long long)B * ((H + Q8_TILE_H - 1) / Q8_TILE_H) * ((W + Q8_TILE_W - 1) / Q8_TILE_W));
}

size_t ssm_quads_bytes(int F, int H, int W) {
    if (F <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)F * (H + 1) * (W + 1) * 16;
}

int ssm_quads_from_u8(const unsigned char* src, long long src_frame_stride, int src_row_stride, int bgr,
                      int F, int H_in, int W_in, int H, int W, int top, int left, void* quads, void* stream) {
    if (F <= 0 || H_in <= 0 || W_in <= 0 || H <= 0 || W <= 0)
        return fail(SSM_ERR_SHAPE, "ssm_quads_from_u8: F, H_in, W_in, H, W must be positive");
    if (top < 0 || left < 0 || top + H_in > H || left + W_in > W)
        return fail(SSM_ERR_SHAPE, "ssm_quads_from_u8: the %d x %d source at (%d, %d) does not fit %d x %d", H_in, W_in, top, left, H, W);
    if ((long long)(H + 1) * (W + 1) > MAX_PLANE) return fail(SSM_ERR_SHAPE, "H*W too large (%d x %d)", H, W);
    if (!src || !quads) return fail(SSM_ERR_NULL, "ssm_quads_from_u8: src or quads is NULL");
    if (((uintptr_t)quads) % 16 != 0) return fail(SSM_ERR_ALIGN, "ssm_quads_from_u8: quads must be 16-byte aligned");
    const long long groups = (long long)F * (H + 1) * ((W + 1 + 3) / 4);
    const unsigned grid = (unsigned)((groups + 255) / 256 < 148ll * 32 ? (groups + 255) / 256 : 148ll * 32);
    quads_from_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_frame_stride, src_row_stride, bgr != 0, H_in, W_in, H, W,
                                                                top, left, (uint4*)quads, groups);
    SSM_LAUNCH_CHECK("ssm_quads_from_u8");
    return SSM_OK;
}

extern "C++" {
template <typename T, int MODE>
static int flow_pack_q8_launch(const ssm_tensor* img6, const void* quads, const ssm_tensor* flow4, const float* t,
                               const ssm_tensor* out16, void* nhwc, int out_dtype, const float* norm6, const float* lut,
                               int B, int N, int H, int W, cudaStream_t s) {
    const Geom g = make_geom(H, W);
    const Norm3 nm = make_norm(norm6);
    const unsigned grid = q8_grid(B, H, W);
    if (!nhwc) {
        flow_pack_fwd_q8_kernel<T, MODE, T, false><<<grid, Q8_THREADS, 0, s>>>(cview<T>(img6), (const uint4*)quads, cview<T>(flow4),
                                                                          t, mview<T>(out16), N, g, nm, lut);
    } else if (out_dtype == SSM_DTYPE_BF16) {
        View<__nv_bfloat16> o; o.p = (__nv_bfloat16*)nhwc; o.sb = (long long)N * 16 * H * W; o.sn = 16ll * H * W; o.sc = 1;
        flow_pack_fwd_q8_kernel<T, MODE, __nv_bfloat16, true><<<grid, Q8_THREADS, 0, s>>>(cview<T>(img6), (const uint4*)quads,
                                                                                     cview<T>(flow4), t, o, N, g, nm, lut);
    } else {
        if constexpr (sizeof(T) != 4) return fail(SSM_ERR_UNSUPPORTED, "ssm_flow_pack_fwd_q8_nhwc: bf16 inputs with fp32 output is not built");
        else {
            View<float> o; o.p = (float*)nhwc; o.sb = (long long)N * 16 * H * W; o.sn = 16ll * H * W; o.sc = 1;
            flow_pack_fwd_q8_kernel<T, MODE, float, true><<<grid, Q8_THREADS, 0, s>>>(cview<T>(img6), (const uint4*)quads, cview<T>(flow4),
                                                                                 t, o, N, g, nm, lut);
        }
    }
    SSM_LAUNCH_CHECK("ssm_flow_pack_fwd_q8");
    return SSM_OK;
}

template <typename T, int MODE, typename TY, bool OUT_U8>
static int fuse_q8_launch(const void* quads, const ssm_tensor* flow4, const ssm_tensor* out5, const float* t,
                          const ssm_tensor* out3, const U8Out& u8, const float* norm6, int B, int N, int H, int W, cudaStream_t s) {
    fuse_fwd_q8_kernel<T, MODE, TY, OUT_U8><<<q8_grid(B, H, W), Q8_THREADS, 0, s>>>(
        (const uint4*)quads, cview<T>(flow4), cview<TY>(out5), t, mview<T>(out3), u8, N, make_geom(H, W), make_norm(norm6));
    SSM_LAUNCH_CHECK("ssm_fuse_flow_fwd_q8");
    return SSM_OK;
}

template <typename T, int MODE>
static int fuse_q8_dispatch(const void* quads, const ssm_tensor* flow4, const ssm_tensor* out5, int out5_dtype, const float* t,
                            const ssm_tensor* out3, const U8Out* u8, const float* norm6, int B, int N, int H, int W, cudaStream_t s) {
    U8Out none = {};
    if (out5_dtype == SSM_DTYPE_BF16)
        return u8 ? fuse_q8_launch<T, MODE, __nv_bfloat16, true>(quads, flow4, out5, t, nullptr, *u8, norm6, B, N, H, W, s)
                  : fuse_q8_launch<T, MODE, __nv_bfloat16, false>(quads, flow4, out5, t, out3, none, norm6, B, N, H, W, s);
    if constexpr (sizeof(T) != 4) return fail(SSM_ERR_UNSUPPORTED, "ssm_fuse_flow_fwd_q8: bf16 flows with an fp32 out5 is not built");
    else
        return u8 ? fuse_q8_launch<T, MODE, float, true>(quads, flow4, out5, t, nullptr, *u8, norm6, B, N, H, W, s)
                  : fuse_q8_launch<T, MODE, float, false>(quads, flow4, out5, t, out3, none, norm6, B, N, H, W, s);
}
}  // extern "C++"

static int flow_pack_q8_impl(const ssm_tensor* img6, const void* quads, const ssm_tensor* flow4, const float* t,
                             const ssm_tensor* out16, void* nhwc, int out_dtype, const float* norm6, const float* lut,
                             int B, int N, int H, int W, int dtype, int coord_mode, void* stream) {
    SSM_TRY(check_q8("ssm_flow_pack_fwd_q8", B, N, H, W, coord_mode, quads, norm6));
    if (dtype != SSM_DTYPE_F32 && dtype != SSM_DTYPE_BF16) return fail(SSM_ERR_DTYPE, "unknown dtype %d", dtype);
    static const ssm_tensor no_frames = {nullptr, 0, 0, 0};
    if (lut) {                                  // pass-through channels from the tables: the planar frames are not read
        if (((uintptr_t)lut) % 4 != 0) return fail(SSM_ERR_ALIGN, "lut is not aligned to 4 bytes");
        img6 = &no_frames;
    } else {
        SSM_TRY(check_tensor(img6, "img6", dtype, true));
        SSM_TRY(check_pairable(img6, "img6", dtype));
    }
    SSM_TRY(check_tensor(flow4, "flow4", dtype, true));
    SSM_TRY(check_pairable(flow4, "flow4", dtype));
    if (!t) return fail(SSM_ERR_NULL, "t is NULL");
    if (nhwc) {
        if (out_dtype != SSM_DTYPE_F32 && out_dtype != SSM_DTYPE_BF16) return fail(SSM_ERR_DTYPE, "unknown out_dtype %d", out_dtype);
        if (((uintptr_t)nhwc) % 32 != 0) return fail(SSM_ERR_ALIGN, "out16_nhwc must be 32-byte aligned");
    } else {
        SSM_TRY(check_tensor(out16, "out16", dtype, true));
        SSM_TRY(check_pairable(out16, "out16", dtype));
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == SSM_DTYPE_F32)
        return coord_mode == SSM_COORD_DIV
            ? flow_pack_q8_launch<float, SSM_COORD_DIV>(img6, quads, flow4, t, out16, nhwc, out_dtype, norm6, lut, B, N, H, W, s)
            : flow_pack_q8_launch<float, SSM_COORD_RCP>(img6, quads, flow4, t, out16, nhwc, out_dtype, norm6, lut, B, N, H, W, s);
    return coord_mode == SSM_COORD_DIV
        ? flow_pack_q8_launch<__nv_bfloat16, SSM_COORD_DIV>(img6, quads, flow4, t, out16, nhwc, out_dtype, norm6, lut, B, N, H, W, s)
        : flow_pack_q8_launch<__nv_bfloat16, SSM_COORD_RCP>(img6, quads, flow4, t, out16, nhwc, out_dtype, norm6, lut, B, N, H, W, s);
}

int ssm_flow_pack_fwd_q8(const ssm_tensor* img6, const void* quads, const ssm_tensor* flow4, const float* t,
                         const ssm_tensor* out16, const float* norm6, int B, int N, int H, int W, int dtype, int coord_mode, void* stream) {
    return flow_pack_q8_impl(img6, quads, flow4, t, out16, nullptr, dtype, norm6, nullptr, B, N, H, W, dtype, coord_mode, stream);
}

int ssm_flow_pack_fwd_q8_nhwc(const ssm_tensor* img6, const void* quads, const ssm_tensor* flow4, const float* t,
                              void* out16_nhwc, int out_dtype, const float* norm6, int B, int N, int H, int W,
                              int dtype, int coord_mode, void* stream) {
    if (!out16_nhwc) return fail(SSM_ERR_NULL, "out16_nhwc is NULL");
    return flow_pack_q8_impl(img6, quads, flow4, t, nullptr, out16_nhwc, out_dtype, norm6, nullptr, B, N, H, W, dtype, coord_mode, stream);
}

int ssm_flow_pack_fwd_q8_lut(const void* quads, const float* lut, const ssm_tensor* flow4, const float* t,
                             const ssm_tensor* out16, const float* norm6, int B, int N, int H, int W, int dtype, int coord_mode, void* stream) {
    if (!lut) return fail(SSM_ERR_NULL, "lut is NULL");
    return flow_pack_q8_impl(nullptr, quads, flow4, t, out16, nullptr, dtype, norm6, lut, B, N, H, W, dtype, coord_mode, stream);
}

static int fuse_q8_impl(const void* quads, const ssm_tensor* flow4, const ssm_tensor* out5, int out5_dtype, const float* t,
                        const ssm_tensor* out3, const U8Out* u8, const float* norm6, int B, int N, int H, int W,
                        int dtype, int coord_mode, void* stream) {
    SSM_TRY(check_q8("ssm_fuse_flow_fwd_q8", B, N, H, W, coord_mode, quads, norm6));
    if (dtype != SSM_DTYPE_F32 && dtype != SSM_DTYPE_BF16) return fail(SSM_ERR_DTYPE, "unknown dtype %d", dtype);
    if (out5_dtype != SSM_DTYPE_F32 && out5_dtype != SSM_DTYPE_BF16) return fail(SSM_ERR_DTYPE, "unknown out5_dtype %d", out5_dtype);
    SSM_TRY(check_tensor(flow4, "flow4", dtype, true));
    SSM_TRY(check_tensor(out5, "out5", out5_dtype, true));
    SSM_TRY(check_pairable(flow4, "flow4", dtype));
    SSM_TRY(check_pairable(out5, "out5", out5_dtype));
    if (!t) return fail(SSM_ERR_NULL, "t is NULL");
    if (!u8) {
        SSM_TRY(check_tensor(out3, "out3", dtype, true));
        SSM_TRY(check_pairable(out3, "out3", dtype));
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == SSM_DTYPE_F32)
        return coord_mode == SSM_COORD_DIV
            ? fuse_q8_dispatch<float, SSM_COORD_DIV>(quads, flow4, out5, out5_dtype, t, out3, u8, norm6, B, N, H, W, s)
            : fuse_q8_dispatch<float, SSM_COORD_RCP>(quads, flow4, out5, out5_dtype, t, out3, u8, norm6, B, N, H, W, s);
    return coord_mode == SSM_COORD_DIV
        ? fuse_q8_dispatch<__nv_bfloat16, SSM_COORD_DIV>(quads, flow4, out5, out5_dtype, t, out3, u8, norm6, B, N, H, W, s)
        : fuse_q8_dispatch<__nv_bfloat16, SSM_COORD_RCP>(quads, flow4, out5, out5_dtype, t, out3, u8, norm6, B, N, H, W, s);
}

int ssm_fuse_flow_fwd_q8(const void* quads, const ssm_tensor* flow4, const ssm_tensor* out5, int out5_dtype, const float* t,
                         const ssm_tensor* out3, const float* norm6, int B, int N, int H, int W, int dtype, int coord_mode, void* stream) {
    return fuse_q8_impl(quads, flow4, out5, out5_dtype, t, out3, nullptr, norm6, B, N, H, W, dtype, coord_mode, stream);
}

int ssm_fuse_flow_fwd_q8_u8(const void* quads, const ssm_tensor* flow4, const ssm_tensor* out5, int out5_dtype, const float* t,
                            unsigned char* dst, long long dst_frame_stride, int dst_row_stride, int top, int left,
                            int H_out, int W_out, const float* mean3, const float* std3, float scale, int bgr, int saturate,
                            const float* norm6, int B, int N, int H, int W, int dtype, int coord_mode, void* stream) {
    if (!dst || !mean3 || !std3) return fail(SSM_ERR_NULL, "ssm_fuse_flow_fwd_q8_u8: dst, mean3 or std3 is NULL");
    if (H_out <= 0 || W_out <= 0 || top < 0 || left < 0 || top + H_out > H || left + W_out > W)
        return fail(SSM_ERR_SHAPE, "ssm_fuse_flow_fwd_q8_u8: the %d x %d crop at (%d, %d) leaves %d x %d", H_out, W_out, top, left, H, W);
    U8Out u8;
    u8.dst = dst; u8.frame_stride = dst_frame_stride; u8.row_stride = dst_row_stride;
    u8.top = top; u8.left = left; u8.H_out = H_out; u8.W_out = W_out; u8.bgr = bgr != 0; u8.saturate = saturate != 0;
    for (int k = 0; k < 3; ++k) { u8.mean[k] = mean3[k]; u8.std[k] = std3[k]; }
    u8.scale = scale;
    return fuse_q8_impl(quads, flow4, out5, out5_dtype, t, nullptr, &u8, norm6, B, N, H, W, dtype, coord_mode, stream);
}

// ---------------------------------------------------------------------------------------------
// Host-buffer entry point: pair-sized chunks, three slots of caller-owned device scratch on three
// streams, so the H2D copy of pair b+1, the kernels of pair b and the D2H copy of pair b-1 overlap.
static size_t host_slot_bytes(int N, int H, int W) {
    const size_t npx = (size_t)H * W;
    size_t floats = (size_t)(6 + 4 + 8 + 5 * N + 16 * N + 3 * N) * npx;   // img, flow, RGBx, out5, in16, out3
    return (floats * sizeof(float) + 64 * sizeof(float) + 255) / 256 * 256;
}

size_t ssm_synthesize_host_scratch_bytes(int B, int N, int H, int W) {
    if (B <= 0 || N <= 0 || H <= 0 || W <= 0) return 0;
    return host_slot_bytes(N, H, W) * (B < 3 ? B : 3);
}

int ssm_synthesize_host(const float* img6_host, const float* flow4_host, const float* out5_host,
                        const float* t_host, float* out3_host, float* in16_host,
                        int B, int N, int H, int W, int coord_mode, void* scratch_v, size_t scratch_bytes) {
    SSM_TRY(check_common(B, N, 16, H, W, SSM_DTYPE_F32, coord_mode));
    if (!img6_host || !flow4_host || !out5_host || !t_host || !out3_host)
        return fail(SSM_ERR_NULL, "ssm_synthesize_host: a required host pointer is NULL");
    if (N > 64) return fail(SSM_ERR_SHAPE, "ssm_synthesize_host: N must be <= 64 (got %d)", N);
    if (!scratch_v || scratch_bytes < ssm_synthesize_host_scratch_bytes(B, N, H, W))
        return fail(SSM_ERR_WORKSPACE, "ssm_synthesize_host: needs %zu bytes of device scratch, got %zu",
                    ssm_synthesize_host_scratch_bytes(B, N, H, W), scratch_v ? scratch_bytes : (size_t)0);
    if (((uintptr_t)scratch_v) % 256 != 0) return fail(SSM_ERR_ALIGN, "scratch must be 256-byte aligned");
    const size_t npx = (size_t)H * W;
    const size_t per_pair = host_slot_bytes(N, H, W);
    const int slots = B < 3 ? B : 3;
    cudaStream_t st[3] = {nullptr, nullptr, nullptr};
    char* scratch = (char*)scratch_v;
    int rc = SSM_OK;
    cudaError_t e;
    for (int i = 0; i < slots && rc == SSM_OK; ++i) {
        e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
        if (e != cudaSuccess) rc = cuda_fail(e, "ssm_synthesize_host cudaStreamCreate");
    }
    for (int b = 0; b < B && rc == SSM_OK; ++b) {
        const int k = b % slots;
        cudaStream_t s = st[k];
        float* d_img = (float*)(scratch + per_pair * k);
        float* d_flow = d_img + 6 * npx;
        float* d_rgbx = d_flow + 4 * npx;
        float* d_out5 = d_rgbx + 8 * npx;
        float* d_in16 = d_out5 + (size_t)5 * N * npx;
        float* d_out3 = d_in16 + (size_t)16 * N * npx;
        float* d_t = d_out3 + (size_t)3 * N * npx;
#define SSM_H(expr) if (rc == SSM_OK && (e = (expr)) != cudaSuccess) rc = cuda_fail(e, "ssm_synthesize_host " #expr)
        SSM_H(cudaMemcpyAsync(d_img, img6_host + (size_t)b * 6 * npx, 6 * npx * sizeof(float), cudaMemcpyHostToDevice, s));
        SSM_H(cudaMemcpyAsync(d_flow, flow4_host + (size_t)b * 4 * npx, 4 * npx * sizeof(float), cudaMemcpyHostToDevice, s));
        SSM_H(cudaMemcpyAsync(d_t, t_host + (size_t)b * N, N * sizeof(float), cudaMemcpyHostToDevice, s));
        SSM_H(cudaMemcpyAsync(d_out5, out5_host + (size_t)b * N * 5 * npx, (size_t)N * 5 * npx * sizeof(float), cudaMemcpyHostToDevice, s));
        if (rc != SSM_OK) break;
        ssm_tensor T_img{d_img, (int64_t)(6 * npx), 0, (int64_t)npx};
        ssm_tensor T_flow{d_flow, (int64_t)(4 * npx), 0, (int64_t)npx};
        ssm_tensor T_out5{d_out5, (int64_t)(5 * N * npx), (int64_t)(5 * npx), (int64_t)npx};
        ssm_tensor T_in16{d_in16, (int64_t)(16 * N * npx), (int64_t)(16 * npx), (int64_t)npx};
        ssm_tensor T_out3{d_out3, (int64_t)(3 * N * npx), (int64_t)(3 * npx), (int64_t)npx};
        const void* rgbx = N >= 2 ? d_rgbx : nullptr;
        if (rgbx) rc = ssm_pack_frames(&T_img, d_rgbx, 1, H, W, SSM_DTYPE_F32, s);
        if (rc == SSM_OK) rc = ssm_flow_pack_fwd(&T_img, rgbx, &T_flow, d_t, &T_in16, 1, N, H, W, SSM_DTYPE_F32, coord_mode, s);
        if (rc == SSM_OK) rc = ssm_fuse_flow_fwd(&T_img, rgbx, &T_flow, &T_out5, d_t, &T_out3, 1, N, H, W, SSM_DTYPE_F32, coord_mode, s);
        SSM_H(cudaMemcpyAsync(out3_host + (size_t)b * N * 3 * npx, d_out3, (size_t)N * 3 * npx * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (in16_host)
            SSM_H(cudaMemcpyAsync(in16_host + (size_t)b * N * 16 * npx, d_in16, (size_t)N * 16 * npx * sizeof(float), cudaMemcpyDeviceToHost, s));
#undef SSM_H
    }
    for (int i = 0; i < slots; ++i)
        if (st[i]) {
            e = cudaStreamSynchronize(st[i]);
            if (e != cudaSuccess && rc == SSM_OK) rc = cuda_fail(e, "ssm_synthesize_host sync");
            cudaStreamDestroy(st[i]);
        }
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Host-buffer entry point for 8-bit frames: uint8 images in, uint8 interpolated images out.  Same 3-slot
// copy/compute pipeline as ssm_synthesize_host; per pair it moves 2 images + flows + the U-Net output up and
// N images down (the fp32 entry moves 6 + 4 + 5N fp32 planes up and 3N fp32 planes down).
struct U8Slot {
    size_t frames, img6, quads, flow, out5, in16, out, t, lut, total;
};
static U8Slot u8_slot(int N, int H_in, int W_in, int H, int W, int out5_dtype) {
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t npx = (size_t)H * W, esz5 = out5_dtype == SSM_DTYPE_F32 ? 4 : 2;
    U8Slot s; size_t o = 0;
    s.frames = o; o += up((size_t)2 * H_in * W_in * 3);
    s.img6 = o;   o += up(6 * npx * 4);
    s.quads = o;  o += up((size_t)2 * (H + 1) * (W + 1) * 16);
    s.flow = o;   o += up(4 * npx * 4);
    s.out5 = o;   o += up((size_t)5 * N * npx * esz5);
    s.in16 = o;   o += up((size_t)16 * N * npx * 4);
    s.out = o;    o += up((size_t)N * H_in * W_in * 3);
    s.t = o;      o += up((size_t)N * 4);
    s.lut = o;    o += up(768 * 4);
    s.total = o;
    return s;
}

size_t ssm_synthesize_host_u8_scratch_bytes(int B, int N, int H_in, int W_in, int H, int W, int out5_dtype) {
    if (B <= 0 || N <= 0 || H_in <= 0 || W_in <= 0 || H <= 0 || W <= 0) return 0;
    return u8_slot(N, H_in, W_in, H, W, out5_dtype).total * (B < 3 ? B : 3);
}

int ssm_synthesize_host_u8(const unsigned char* frames_host, int bgr, const float* flow4_host, const void* out5_host,
                           int out5_dtype, const float* t_host, unsigned char* out_host, const float* lut_host,
                           const float* norm6, const float* mean3, const float* std3, int saturate,
                           int B, int N, int H_in, int W_in, int H, int W, int top, int left, int coord_mode,
                           void* scratch_v, size_t scratch_bytes) {
    SSM_TRY(check_common(B, N, 16, H, W, SSM_DTYPE_F32, coord_mode));
    if (!frames_host || !flow4_host || !out5_host || !t_host || !out_host || !lut_host || !norm6 || !mean3 || !std3)
        return fail(SSM_ERR_NULL, "ssm_synthesize_host_u8: a required host pointer is NULL");
    if (out5_dtype != SSM_DTYPE_F32 && out5_dtype != SSM_DTYPE_BF16) return fail(SSM_ERR_DTYPE, "unknown out5_dtype %d", out5_dtype);
    if (H_in <= 0 || W_in <= 0 || top < 0 || left < 0 || top + H_in > H || left + W_in > W)
        return fail(SSM_ERR_SHAPE, "ssm_synthesize_host_u8: the %d x %d images at (%d, %d) do not fit %d x %d", H_in, W_in, top, left, H, W);
    if (W % 4 != 0) return fail(SSM_ERR_SHAPE, "ssm_synthesize_host_u8: padded width must be a multiple of 4 (got %d)", W);
    if (!scratch_v || scratch_bytes < ssm_synthesize_host_u8_scratch_bytes(B, N, H_in, W_in, H, W, out5_dtype))
        return fail(SSM_ERR_WORKSPACE, "ssm_synthesize_host_u8: needs %zu bytes of device scratch, got %zu",
                    ssm_synthesize_host_u8_scratch_bytes(B, N, H_in, W_in, H, W, out5_dtype), scratch_v ? scratch_bytes : (size_t)0);
    if (((uintptr_t)scratch_v) % 256 != 0) return fail(SSM_ERR_ALIGN, "scratch must be 256-byte aligned");
    const U8Slot L = u8_slot(N, H_in, W_in, H, W, out5_dtype);
    const size_t npx = (size_t)H * W, esz5 = out5_dtype == SSM_DTYPE_F32 ? 4 : 2;
    const size_t img_bytes = (size_t)H_in * W_in * 3;
    const int slots = B < 3 ? B : 3;
    cudaStream_t st[3] = {nullptr, nullptr, nullptr};
    int rc = SSM_OK;
    cudaError_t e;
    for (int i = 0; i < slots && rc == SSM_OK; ++i) {
        e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
        if (e != cudaSuccess) rc = cuda_fail(e, "ssm_synthesize_host_u8 cudaStreamCreate");
    }
    const float pad3[3] = {lut_host[0], lut_host[256], lut_host[512]};      // byte 0, normalised (pad before normalising)
    for (int b = 0; b < B && rc == SSM_OK; ++b) {
        const int k = b % slots;
        cudaStream_t s = st[k];
        char* base = (char*)scratch_v + L.total * k;
        unsigned char* d_frames = (unsigned char*)(base + L.frames);
        float* d_img = (float*)(base + L.img6);
        void* d_quads = base + L.quads;
        float* d_flow = (float*)(base + L.flow);
        void* d_out5 = base + L.out5;
        float* d_in16 = (float*)(base + L.in16);
        unsigned char* d_out = (unsigned char*)(base + L.out);
        float* d_t = (float*)(base + L.t);
        float* d_lut = (float*)(base + L.lut);
#define SSM_H(expr) if (rc == SSM_OK && (e = (expr)) != cudaSuccess) rc = cuda_fail(e, "ssm_synthesize_host_u8 " #expr)
        SSM_H(cudaMemcpyAsync(d_frames, frames_host + (size_t)b * 2 * img_bytes, 2 * img_bytes, cudaMemcpyHostToDevice, s));
        SSM_H(cudaMemcpyAsync(d_flow, flow4_host + (size_t)b * 4 * npx, 4 * npx * sizeof(float), cudaMemcpyHostToDevice, s));
        SSM_H(cudaMemcpyAsync(d_t, t_host + (size_t)b * N, N * sizeof(float), cudaMemcpyHostToDevice, s));
        if (b < slots) SSM_H(cudaMemcpyAsync(d_lut, lut_host, 768 * sizeof(float), cudaMemcpyHostToDevice, s));
        SSM_H(cudaMemcpyAsync(d_out5, (const char*)out5_host + (size_t)b * N * 5 * npx * esz5, (size_t)N * 5 * npx * esz5, cudaMemcpyHostToDevice, s));
        if (rc != SSM_OK) break;
        ssm_tensor T_img{d_img, (int64_t)(3 * npx), 0, (int64_t)npx};            // as F = 2 frames of 3 planes
        ssm_tensor T_img6{d_img, (int64_t)(6 * npx), 0, (int64_t)npx};
        ssm_tensor T_flow{d_flow, (int64_t)(4 * npx), 0, (int64_t)npx};
        ssm_tensor T_out5{d_out5, (int64_t)(5 * N * npx), (int64_t)(5 * npx), (int64_t)npx};
        ssm_tensor T_in16{d_in16, (int64_t)(16 * N * npx), (int64_t)(16 * npx), (int64_t)npx};
        rc = ssm_frames_from_u8(d_frames, (long long)img_bytes, W_in * 3, bgr, 2, H_in, W_in, H, W, top, left, d_lut, pad3, &T_img, nullptr,
                                SSM_DTYPE_F32, s);
        if (rc == SSM_OK) rc = ssm_quads_from_u8(d_frames, (long long)img_bytes, W_in * 3, bgr, 2, H_in, W_in, H, W, top, left, d_quads, s);
        if (rc == SSM_OK) rc = ssm_flow_pack_fwd_q8(&T_img6, d_quads, &T_flow, d_t, &T_in16, norm6, 1, N, H, W, SSM_DTYPE_F32, coord_mode, s);
        if (rc == SSM_OK) rc = ssm_fuse_flow_fwd_q8_u8(d_quads, &T_flow, &T_out5, out5_dtype, d_t, d_out, (long long)img_bytes, W_in * 3, top, left,
                                                       H_in, W_in, mean3, std3, 255.0f, bgr, saturate, norm6, 1, N, H, W, SSM_DTYPE_F32, coord_mode, s);
        SSM_H(cudaMemcpyAsync(out_host + (size_t)b * N * img_bytes, d_out, (size_t)N * img_bytes, cudaMemcpyDeviceToHost, s));
#undef SSM_H
    }
    for (int i = 0; i < slots; ++i)
        if (st[i]) {
            e = cudaStreamSynchronize(st[i]);
            if (e != cudaSuccess && rc == SSM_OK) rc = cuda_fail(e, "ssm_synthesize_host_u8 sync");
            cudaStreamDestroy(st[i]);
        }
    return rc;
}

}  // extern "C"
