// tif_warp.cu -- flow-driven remaps of libtif_b200.so (sm_100a): flow_warp.warp_backward (flow_warp.py:54-88),
// flow_warp.flow_warp_meshgrid (:15-51) and flow_postproc.erp_of_wraparound (flow_postproc.py:17-44).
//
//   k_warp_rgb_fast   uint8 RGB image, scipy 'wrap' boundary: the evaluation remap of BASELINE.json config #3.
//                     One thread = 4 horizontally adjacent pixels, a warp = a 32 x 4 pixel tile; coordinates stay in
//                     integers + a 23-bit fraction (no float64, no XU-pipe conversions), the taps are fetched as
//                     aligned words and blended in 2^24 fixed point; a sample whose rounding the fast arithmetic
//                     cannot decide is redone by the float64 reference expression, so the uint8 output is scipy's.
//   k_warp_backward   every other case (any channel count, float images, 'constant' boundary): float64 throughout.
#include <string.h>

#include "tif_bilinear.cuh"

// ------------------------------------------------------------------------------------------------
// generic kernel: one thread per pixel, float64 coordinates exactly as the reference computes them
// ------------------------------------------------------------------------------------------------
// flow_warp.py:75-86: (x, y) = mesh grid + flow in float64, scipy boundary rule, base tap clamped to n-2
template <int MODE>
__device__ __forceinline__ bool warp_coords_f64(double x, double y, int h, int w, double& x0, double& y0, double& fx,
                                                double& fy) {
    if (MODE == TIF_WARP_CONSTANT255) {
        if (x < 0.0 || x > (double)(w - 1) || y < 0.0 || y > (double)(h - 1)) return false;
    } else {
        x = scipy_wrap(x, w);
        y = scipy_wrap(y, h);
    }
    x0 = floor(x);
    y0 = floor(y);
    if (x0 > (double)(w - 2)) x0 = (double)(w - 2);
    if (y0 > (double)(h - 2)) y0 = (double)(h - 2);
    if (x0 < 0.0) x0 = 0.0;
    if (y0 < 0.0) y0 = 0.0;
    fx = x - x0;
    fy = y - y0;
    return true;
}

template <typename ImgT, typename FlowT, int MODE>
__global__ void __launch_bounds__(256)
k_warp_backward(const ImgT* __restrict__ image, const FlowT* __restrict__ flow, ImgT* __restrict__ out, int batch,
                int h, int w, int channels) {
    const int64_t n_pix = (int64_t)h * w;
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (pix >= n_pix) return;
    const int r = (int)(pix / w), c = (int)(pix - (int64_t)r * w);
    const size_t gi = (size_t)b * n_pix + pix;
    ImgT* o = out + gi * channels;
    double x0, y0, fx, fy;
    if (!warp_coords_f64<MODE>((double)c + (double)flow[2 * gi], (double)r + (double)flow[2 * gi + 1], h, w, x0, y0, fx, fy)) {
        for (int ch = 0; ch < channels; ++ch) o[ch] = (ImgT)255;
        return;
    }
    const ImgT* r0 = image + ((size_t)b * n_pix + (size_t)y0 * w + (size_t)x0) * channels;
    const ImgT* r1 = r0 + (size_t)w * channels;
    if (sizeof(ImgT) == 1) {
        const float fxf = (float)fx, fyf = (float)fy;
        const float wx0 = 1.0f - fxf, wy0 = 1.0f - fyf;
        for (int ch = 0; ch < channels; ++ch) {
            const float p00 = (float)__ldg(r0 + ch), p01 = (float)__ldg(r0 + channels + ch);
            const float p10 = (float)__ldg(r1 + ch), p11 = (float)__ldg(r1 + channels + ch);
            const float t = bilinear_f32(p00, p01, p10, p11, wx0, fxf, wy0, fyf);
            o[ch] = (ImgT)round_u8_guarded(t, p00, p01, p10, p11, fx, fy);
        }
    } else {
        // float images: scipy accumulates in double and casts to the image's dtype
        for (int ch = 0; ch < channels; ++ch) {
            const double t = bilinear_f64_exact((double)__ldg(r0 + ch), (double)__ldg(r0 + channels + ch),
                                                (double)__ldg(r1 + ch), (double)__ldg(r1 + channels + ch), fx, fy);
            o[ch] = (ImgT)t;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fast kernel: uint8 RGB, 'wrap' boundary
// ------------------------------------------------------------------------------------------------
// Error budget of the fast arithmetic, in units of 2^-24 of a grey level: the fractions are kept to 23 bits
// (|error| <= 2^-24 + 2^-25 per axis -> 255 * 2 * 1.5 = 765 units), the four weights are rounded to 2^-24 and sum to
// exactly 2^24 (<= 4 * 255 * 0.5 = 510 units): 1275 units in total.  A sample whose value + 0.5 lies within 2048 units
// (1.2e-4) of an integer is therefore re-evaluated by the float64 reference expression (7e-4 of the pixels).
#define TIF_WARP_GUARD 0x00000800u
#define TIF_WARP_MAX_FLOW 1048576.0f      // |flow| above this (or NaN) takes the float64 coordinates

__device__ __forceinline__ uint32_t warp_fix_bilinear(uint32_t p00, uint32_t p01, uint32_t p10, uint32_t p11, uint32_t w00,
                                                      uint32_t w01, uint32_t w10, uint32_t w11) {
    return p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + (TIF_FIX_HALF + TIF_WARP_GUARD);
}

__device__ __forceinline__ bool warp_fix_undecided(uint32_t acc) {
    return (acc & (0x01000000u - 2u * TIF_WARP_GUARD)) == 0u;
}

// the float64 reference evaluation of one RGB pixel (rare: undecided roundings, image tail)
template <typename FlowT>
__device__ __noinline__ uint32_t warp_pixel_exact(const uint8_t* __restrict__ image_b, FlowT u, FlowT v, int r, int c, int h,
                                                  int w) {
    double x0, y0, fx, fy;
    warp_coords_f64<TIF_WARP_WRAP>((double)c + (double)u, (double)r + (double)v, h, w, x0, y0, fx, fy);
    const uint8_t* r0 = image_b + ((size_t)y0 * w + (size_t)x0) * 3;
    const uint8_t* r1 = r0 + (size_t)w * 3;
    uint32_t px = 0xff000000u;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const double t = bilinear_f64_exact((double)__ldg(r0 + ch), (double)__ldg(r0 + 3 + ch), (double)__ldg(r1 + ch),
                                            (double)__ldg(r1 + 3 + ch), fx, fy);
        px |= (uint32_t)scipy_round_u8(t) << (8 * ch);
    }
    return px;
}

// integer base tap + 23-bit fraction of one axis.  float32 flow: floor and fraction of the flow itself, added to the
// integer pixel position (no float64); anything that leaves [0, n-1] (scipy wraps with period n-1 there), huge or
// NaN takes the float64 route.  float64 flow: always the float64 route.
template <typename FlowT>
__device__ __forceinline__ void warp_axis(FlowT d, int pos, int n, int& i0, uint32_t& q) {
    if (sizeof(FlowT) == 4) {
        const float df = (float)d;
        const float fl = floorf(df);
        const int xi = pos + (int)fl;
        // (df - fl) + 1 rounds the fraction to 23 bits; the integer difference to 1.0f is the fraction * 2^23 (<= 2^23)
        const uint32_t frac = (uint32_t)(__float_as_int((df - fl) + 1.0f) - 0x3f800000);
        if (fabsf(df) < TIF_WARP_MAX_FLOW && xi >= 0 && xi < n - 1) {
            i0 = xi;
            q = frac;
            return;
        }
    }
    double x = scipy_wrap((double)pos + (double)d, n);
    double x0 = floor(x);
    if (x0 > (double)(n - 2)) x0 = (double)(n - 2);
    if (x0 < 0.0) x0 = 0.0;
    i0 = (int)x0;
    q = (uint32_t)((x - x0) * 8388608.0 + 0.5);
}

#ifndef TIF_WARP_MIN_BLOCKS
#define TIF_WARP_MIN_BLOCKS 8
#endif
template <typename FlowT>
__global__ void __launch_bounds__(128, TIF_WARP_MIN_BLOCKS)
k_warp_rgb_fast(const uint8_t* __restrict__ image, const FlowT* __restrict__ flow, uint8_t* __restrict__ out, int batch,
                int h, int w) {
    // warp -> 32 x 4 pixel tile, lane -> 4 adjacent pixels of one row
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_x = (w + 31) >> 5;
    const int ty = (int)(warp / tiles_x), tx = (int)(warp - (int64_t)ty * tiles_x);
    const int r = ty * 4 + (lane >> 3), c0 = tx * 32 + (lane & 7) * 4;
    if (r >= h || c0 >= w) return;          // w % 4 == 0: a quad is inside or outside as a whole
    const int b = blockIdx.y;
    const size_t n_pix = (size_t)h * w;
    const size_t image_bytes = n_pix * 3;
    const uint8_t* __restrict__ img = image + (size_t)b * image_bytes;
    const size_t gi = (size_t)b * n_pix + (size_t)r * w + c0;
    FlowT fu[4], fv[4];
    if (sizeof(FlowT) == 4) {
        const float4* f4 = reinterpret_cast<const float4*>(flow + 2 * gi);
        const float4 a = __ldg(f4), bq = __ldg(f4 + 1);
        fu[0] = a.x; fv[0] = a.y; fu[1] = a.z; fv[1] = a.w;
        fu[2] = bq.x; fv[2] = bq.y; fu[3] = bq.z; fv[3] = bq.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            fu[j] = flow[2 * (gi + j)];
            fv[j] = flow[2 * (gi + j) + 1];
        }
    }
    const uint32_t row_bytes = (uint32_t)w * 3u;
    uint32_t off[4], qx[4], qy[4];
    RawPair t0[4], t1[4];
    bool tail[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int x0, y0;
        warp_axis<FlowT>(fu[j], c0 + j, w, x0, qx[j]);
        warp_axis<FlowT>(fv[j], r, h, y0, qy[j]);
        off[j] = (uint32_t)y0 * row_bytes + (uint32_t)x0 * 3u;
        // the aligned 12-byte windows of the two tap rows may reach 9 bytes past the last tap: only next to the end
        // of the whole buffer (last image) could that leave it
        tail[j] = (b == batch - 1) && ((size_t)off[j] + row_bytes + 16 > image_bytes);
        if (!tail[j]) {
            const uint8_t* p0 = img + off[j];
            t0[j] = load_raw_pair(p0, (unsigned)(reinterpret_cast<uintptr_t>(p0) & 3));
            t1[j] = load_raw_pair(p0 + row_bytes, (unsigned)(reinterpret_cast<uintptr_t>(p0 + row_bytes) & 3));
        }
    }
    uint32_t px[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (tail[j]) {
            px[j] = warp_pixel_exact<FlowT>(img, fu[j], fv[j], r, c0 + j, h, w);
            continue;
        }
        const uint8_t* p0 = img + off[j];
        uint32_t lo0, hi0, lo1, hi1;
        align_pair(t0[j], (unsigned)(reinterpret_cast<uintptr_t>(p0) & 3), lo0, hi0);
        align_pair(t1[j], (unsigned)(reinterpret_cast<uintptr_t>(p0 + row_bytes) & 3), lo1, hi1);
        // weights in 2^-24 from the 23-bit fractions X, Y: A = round(X Y / 2^22);  w11 = A, w01 = 2X - A, w10 = 2Y - A,
        // w00 = 2^24 - 2X - 2Y + A (they sum to 2^24 exactly)
        const uint32_t a = (uint32_t)(((unsigned long long)qx[j] * qy[j] + (1ull << 21)) >> 22);
        const uint32_t w11 = a, w01 = 2u * qx[j] - a, w10 = 2u * qy[j] - a, w00 = 0x01000000u - 2u * qx[j] - 2u * qy[j] + a;
        const uint32_t ar = warp_fix_bilinear(TIF_BYTE(lo0, 0), TIF_BYTE(lo0, 3), TIF_BYTE(lo1, 0), TIF_BYTE(lo1, 3), w00, w01, w10, w11);
        const uint32_t ag = warp_fix_bilinear(TIF_BYTE(lo0, 1), TIF_BYTE(hi0, 0), TIF_BYTE(lo1, 1), TIF_BYTE(hi1, 0), w00, w01, w10, w11);
        const uint32_t ab = warp_fix_bilinear(TIF_BYTE(lo0, 2), TIF_BYTE(hi0, 1), TIF_BYTE(lo1, 2), TIF_BYTE(hi1, 1), w00, w01, w10, w11);
        if (warp_fix_undecided(ar) | warp_fix_undecided(ag) | warp_fix_undecided(ab))
            px[j] = warp_pixel_exact<FlowT>(img, fu[j], fv[j], r, c0 + j, h, w);
        else
            px[j] = __byte_perm(__byte_perm(ar, ag, 0x0073u), ab, 0x0710u);          // R | G << 8 | B << 16
    }
    // 4 x RGB = 12 bytes = three aligned words
    uint32_t* o = reinterpret_cast<uint32_t*>(out + gi * 3);
    o[0] = __byte_perm(px[0], px[1], 0x4210u);
    o[1] = __byte_perm(px[1], px[2], 0x5421u);
    o[2] = __byte_perm(px[2], px[3], 0x6542u);
}

template <typename ImgT>
static int launch_warp(const ImgT* image, const void* flow, int flow_dtype, int batch, int h, int w, int channels,
                       int mode, ImgT* out, void* stream_v) {
    if (!image || !flow || !out || batch <= 0 || h < 2 || w < 2 || channels <= 0) {
        tif_set_error("tif_warp_backward: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    if (flow_dtype != TIF_DTYPE_F32 && flow_dtype != TIF_DTYPE_F64) {
        tif_set_error("tif_warp_backward: unknown flow dtype %d", flow_dtype);
        return TIF_ERR_INVALID_ARG;
    }
    if (mode != TIF_WARP_WRAP && mode != TIF_WARP_CONSTANT255) {
        tif_set_error("tif_warp_backward: unknown mode %d", mode);
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (sizeof(ImgT) == 1 && channels == 3 && mode == TIF_WARP_WRAP && w % 4 == 0 && (size_t)h * w * 3 < ((size_t)1 << 32) &&
        reinterpret_cast<uintptr_t>(image) % 4 == 0 && reinterpret_cast<uintptr_t>(out) % 4 == 0 &&
        reinterpret_cast<uintptr_t>(flow) % 16 == 0) {
        const int64_t warps = (int64_t)((w + 31) / 32) * ((h + 3) / 4);
        dim3 grid((unsigned)((warps * 32 + 127) / 128), (unsigned)batch);
        if (flow_dtype == TIF_DTYPE_F32)
            k_warp_rgb_fast<float><<<grid, 128, 0, stream>>>((const uint8_t*)image, (const float*)flow, (uint8_t*)out, batch, h, w);
        else
            k_warp_rgb_fast<double><<<grid, 128, 0, stream>>>((const uint8_t*)image, (const double*)flow, (uint8_t*)out, batch, h, w);
        TIF_CUDA_CHECK(cudaGetLastError());
        return TIF_OK;
    }
    const int threads = 256;
    dim3 grid((unsigned)(((int64_t)h * w + threads - 1) / threads), (unsigned)batch);
    if (flow_dtype == TIF_DTYPE_F32 && mode == TIF_WARP_WRAP)
        k_warp_backward<ImgT, float, TIF_WARP_WRAP><<<grid, threads, 0, stream>>>(image, (const float*)flow, out, batch, h, w, channels);
    else if (flow_dtype == TIF_DTYPE_F32)
        k_warp_backward<ImgT, float, TIF_WARP_CONSTANT255><<<grid, threads, 0, stream>>>(image, (const float*)flow, out, batch, h, w, channels);
    else if (mode == TIF_WARP_WRAP)
        k_warp_backward<ImgT, double, TIF_WARP_WRAP><<<grid, threads, 0, stream>>>(image, (const double*)flow, out, batch, h, w, channels);
    else
        k_warp_backward<ImgT, double, TIF_WARP_CONSTANT255><<<grid, threads, 0, stream>>>(image, (const double*)flow, out, batch, h, w, channels);
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

extern "C" int tif_warp_backward_u8(const uint8_t* image, const void* flow, int flow_dtype, int batch, int h, int w,
                                    int channels, int mode, uint8_t* out, void* stream) {
    return launch_warp<uint8_t>(image, flow, flow_dtype, batch, h, w, channels, mode, out, stream);
}

extern "C" int tif_warp_backward_f32(const float* image, const void* flow, int flow_dtype, int batch, int h, int w,
                                     int channels, int mode, float* out, void* stream) {
    return launch_warp<float>(image, flow, flow_dtype, batch, h, w, channels, mode, out, stream);
}

extern "C" int tif_warp_backward_f64(const double* image, const void* flow, int flow_dtype, int batch, int h, int w,
                                     int channels, int mode, double* out, void* stream) {
    return launch_warp<double>(image, flow, flow_dtype, batch, h, w, channels, mode, out, stream);
}

// ------------------------------------------------------------------------------------------------
// flow_warp.flow_warp_meshgrid / flow_postproc.erp_of_wraparound
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_meshgrid(const T* __restrict__ fu, const T* __restrict__ fv, T* __restrict__ out, int batch, int h, int w) {
    const int64_t n_pix = (int64_t)h * w;
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (pix >= n_pix) return;
    const int r = (int)(pix / w), c = (int)(pix - (int64_t)r * w);
    double eu = (double)c + (double)fu[(size_t)b * n_pix + pix];
    double ev = (double)r + (double)fv[(size_t)b * n_pix + pix];
    // flow_warp.py:41-49: one wrap each way, not a modulo
    if (eu >= (double)w - 0.5) eu -= (double)w;
    if (eu < -0.5) eu += (double)w;
    if (ev >= (double)h - 0.5) ev -= (double)h;
    if (ev < -0.5) ev += (double)h;
    out[((size_t)b * 2 + 0) * n_pix + pix] = (T)eu;
    out[((size_t)b * 2 + 1) * n_pix + pix] = (T)ev;
}

extern "C" int tif_flow_warp_meshgrid(const void* flow_u, const void* flow_v, int dtype, int batch, int h, int w,
                                      void* out, void* stream_v) {
    if (!flow_u || !flow_v || !out || batch <= 0 || h <= 0 || w <= 0) {
        tif_set_error("tif_flow_warp_meshgrid: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    dim3 grid((unsigned)(((int64_t)h * w + 255) / 256), (unsigned)batch);
    if (dtype == TIF_DTYPE_F32)
        k_meshgrid<float><<<grid, 256, 0, stream>>>((const float*)flow_u, (const float*)flow_v, (float*)out, batch, h, w);
    else if (dtype == TIF_DTYPE_F64)
        k_meshgrid<double><<<grid, 256, 0, stream>>>((const double*)flow_u, (const double*)flow_v, (double*)out, batch, h, w);
    else
        return TIF_ERR_INVALID_ARG;
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_wraparound(T* __restrict__ flow, int batch, int h, int w, double thr) {
    const int64_t n_pix = (int64_t)h * w;
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (pix >= n_pix) return;
    const int c = (int)(pix % w);
    const int split = (int)((double)w - 1.0 - thr);
    T* p = flow + ((size_t)b * n_pix + pix) * 2;
    const double u = (double)p[0];
    if (c < split) {
        if (u > thr) p[0] = (T)(u - (double)w);
    } else if (c < w - 1) {
        if (u < -thr) p[0] = (T)(u + (double)w);
    }
}

extern "C" int tif_erp_of_wraparound(void* erp_flow, int dtype, int batch, int h, int w, double u_threshold,
                                     void* stream_v) {
    if (!erp_flow || batch <= 0 || h <= 0 || w <= 0) {
        tif_set_error("tif_erp_of_wraparound: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    dim3 grid((unsigned)(((int64_t)h * w + 255) / 256), (unsigned)batch);
    if (dtype == TIF_DTYPE_F32)
        k_wraparound<float><<<grid, 256, 0, stream>>>((float*)erp_flow, batch, h, w, u_threshold);
    else if (dtype == TIF_DTYPE_F64)
        k_wraparound<double><<<grid, 256, 0, stream>>>((double*)erp_flow, batch, h, w, u_threshold);
    else
        return TIF_ERR_INVALID_ARG;
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

// ------------------------------------------------------------------------------------------------
// compact forms of a tangent stack for the trip back to the host
// ------------------------------------------------------------------------------------------------
// The stack is RGBA with a constant alpha of 255 (full_face_image) -- a quarter of the bytes that cross PCIe say
// nothing -- and its consumer, the per-face estimator, starts with cv2.cvtColor(img[:, :, :3], COLOR_BGR2GRAY)
// (flow_estimate.py:31-47): OpenCV's 8-bit path is (3735 c0 + 19235 c1 + 9798 c2 + 2^14) >> 15, reproduced exactly.
// One thread packs 4 pixels: 16 bytes in, 12 (RGB) or 4 (grey) out, all aligned words.
template <int FORMAT>
__global__ void __launch_bounds__(256)
k_pack_tangent(const uint4* __restrict__ rgba4, int64_t n_quads, uint32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_quads) return;
    const uint4 q = __ldg(rgba4 + i);
    if (FORMAT == TIF_PACK_RGB) {
        out[3 * i] = __byte_perm(q.x, q.y, 0x4210u);
        out[3 * i + 1] = __byte_perm(q.y, q.z, 0x5421u);
        out[3 * i + 2] = __byte_perm(q.z, q.w, 0x6542u);
    } else {
        const uint32_t px[4] = {q.x, q.y, q.z, q.w};
        uint32_t g = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t y = ((px[j] & 0xffu) * 3735u + ((px[j] >> 8) & 0xffu) * 19235u + ((px[j] >> 16) & 0xffu) * 9798u + 16384u) >> 15;
            g |= y << (8 * j);
        }
        out[i] = g;
    }
}

extern "C" int tif_pack_tangent_u8(const uint8_t* tangent_rgba, int64_t n_pixels, int format, uint8_t* out, void* stream_v) {
    if (!tangent_rgba || !out || n_pixels <= 0 || n_pixels % 4 != 0 || reinterpret_cast<uintptr_t>(tangent_rgba) % 16 != 0 ||
        reinterpret_cast<uintptr_t>(out) % 4 != 0) {
        tif_set_error("tif_pack_tangent_u8: invalid argument (n_pixels must be a multiple of 4, buffers 16 / 4 byte aligned)");
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int64_t n_quads = n_pixels / 4;
    const unsigned blocks = (unsigned)((n_quads + 255) / 256);
    if (format == TIF_PACK_RGB)
        k_pack_tangent<TIF_PACK_RGB><<<blocks, 256, 0, stream>>>((const uint4*)tangent_rgba, n_quads, (uint32_t*)out);
    else if (format == TIF_PACK_GRAY_CV)
        k_pack_tangent<TIF_PACK_GRAY_CV><<<blocks, 256, 0, stream>>>((const uint4*)tangent_rgba, n_quads, (uint32_t*)out);
    else {
        tif_set_error("tif_pack_tangent_u8: unknown format %d", format);
        return TIF_ERR_INVALID_ARG;
    }
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

// ------------------------------------------------------------------------------------------------
// peer-mapped buffers: collecting the stitched flows of all ranks on one GPU
// ------------------------------------------------------------------------------------------------
// One process per GPU.  The collecting rank allocates the [n_items, H, W, 2] buffer with tif_ipc_alloc and exports
// its CUDA IPC handle; every other rank opens it (cudaIpcOpenMemHandle maps the peer's memory over NVLink / NVSwitch
// and enables peer access) and hands the address of its own slice to tif_ico2erp_flow_f32 as `erp_flow`: the stores
// of the blend kernel's epilogue then travel straight into the collecting GPU's HBM -- no staging buffer, no extra
// kernel, no receive side.  tif_ipc_copy_async is the alternative for a caller that wants the result locally too.
extern "C" int tif_ipc_alloc(void** dev_ptr, int64_t bytes) {
    if (!dev_ptr || bytes <= 0) {
        tif_set_error("tif_ipc_alloc: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    *dev_ptr = nullptr;
    const cudaError_t e = cudaMalloc(dev_ptr, (size_t)bytes);       // a plain allocation: pool memory cannot be exported
    if (e != cudaSuccess) {
        tif_set_error("tif_ipc_alloc: cudaMalloc(%lld bytes) failed: %s", (long long)bytes, cudaGetErrorString(e));
        return TIF_ERR_ALLOC;
    }
    return TIF_OK;
}

extern "C" int tif_ipc_free(void* dev_ptr) {
    if (dev_ptr) TIF_CUDA_CHECK(cudaFree(dev_ptr));
    return TIF_OK;
}

static_assert(sizeof(cudaIpcMemHandle_t) == TIF_IPC_HANDLE_BYTES, "TIF_IPC_HANDLE_BYTES must match cudaIpcMemHandle_t");

extern "C" int tif_ipc_export(void* dev_ptr, unsigned char* handle) {
    if (!dev_ptr || !handle) {
        tif_set_error("tif_ipc_export: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    cudaIpcMemHandle_t h;
    TIF_CUDA_CHECK(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle, &h, sizeof(h));
    return TIF_OK;
}

extern "C" int tif_ipc_open(const unsigned char* handle, void** dev_ptr) {
    if (!handle || !dev_ptr) {
        tif_set_error("tif_ipc_open: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    *dev_ptr = nullptr;
    TIF_CUDA_CHECK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TIF_OK;
}

extern "C" int tif_ipc_close(void* dev_ptr) {
    if (dev_ptr) TIF_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
    return TIF_OK;
}

extern "C" int tif_ipc_copy_async(void* dst, const void* src, int64_t bytes, void* stream) {
    if (!dst || !src || bytes < 0) {
        tif_set_error("tif_ipc_copy_async: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    TIF_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return TIF_OK;
}
