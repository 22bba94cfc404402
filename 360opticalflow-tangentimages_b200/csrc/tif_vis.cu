// tif_vis.cu -- Middlebury colour-wheel rendering of a flow field (flow_vis.py:77-201; row f4 of SURVEY.md section 8).
// Per-pixel work only: radius, angle, two colour-wheel rows, interpolation, floor(255 c).  float64 throughout, in the
// reference's operation order (the 55 x 3 wheel comes from the host, flow_vis.make_colorwheel).
#include "tif_common.cuh"

#define TIF_WHEEL_ROWS 55

// max over the field of sqrt(u^2 + v^2) after the optional clip (flow_vis.py:147-156): non-negative doubles order
// like their bit patterns, so one 64-bit atomicMax per block does it
template <typename T>
__global__ void __launch_bounds__(256)
k_flow_rad_max(const T* __restrict__ flow, int64_t n, int use_clip, double lo, double hi, unsigned long long* __restrict__ out) {
    double best = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double u = (double)flow[2 * i], v = (double)flow[2 * i + 1];
        if (use_clip) {
            u = fmin(fmax(u, lo), hi);
            v = fmin(fmax(v, lo), hi);
        }
        const double r = sqrt(u * u + v * v);
        if (r > best) best = r;             // NaN never wins, like a nan-free field; np.max would propagate it
    }
    for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_down_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(best));
}

template <typename T>
__global__ void __launch_bounds__(256)
k_flow_to_color(const T* __restrict__ flow, int64_t n, int use_clip, double lo, double hi, double norm,
                const double* __restrict__ wheel /*[55][3]*/, int to_bgr, uint8_t* __restrict__ out) {
    __shared__ double sh_wheel[TIF_WHEEL_ROWS * 3];
    for (int i = threadIdx.x; i < TIF_WHEEL_ROWS * 3; i += blockDim.x) sh_wheel[i] = wheel[i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double u = (double)flow[2 * i], v = (double)flow[2 * i + 1];
    if (use_clip) {
        u = fmin(fmax(u, lo), hi);
        v = fmin(fmax(v, lo), hi);
    }
    if (norm > 0.0) {                       // flow_vis.py:157-162: u / (rad_max + 1e-5)
        u = u / norm;
        v = v / norm;
    }
    // flow_uv_to_colors, flow_vis.py:101-121
    const double rad = sqrt(u * u + v * v);
    const double angle = atan2(-v, -u) / TIF_PI;
    const double row = (angle + 1.0) / 2.0 * (double)(TIF_WHEEL_ROWS - 1);
    const double fl = floor(row);
    const int k0 = (int)fl;
    int k1 = k0 + 1;
    if (k1 == TIF_WHEEL_ROWS) k1 = 0;
    const double ratio = row - fl;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const double c0 = sh_wheel[k0 * 3 + ch] / 255.0, c1 = sh_wheel[k1 * 3 + ch] / 255.0;
        double color = (1.0 - ratio) * c0 + ratio * c1;
        if (rad <= 1.0) color = 1.0 - rad * (1.0 - color);
        else color = color * 0.75;
        out[3 * i + (to_bgr ? 2 - ch : ch)] = (uint8_t)floor(255.0 * color);
    }
}

extern "C" int tif_flow_rad_max(const void* flow, int dtype, int64_t n_pixels, int use_clip, double clip_lo, double clip_hi,
                                double* out /*device, 1*/, void* stream_v) {
    if (!flow || !out || n_pixels <= 0 || (dtype != TIF_DTYPE_F32 && dtype != TIF_DTYPE_F64)) {
        tif_set_error("tif_flow_rad_max: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    TIF_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(double), stream));
    const unsigned blocks = (unsigned)((n_pixels + 255) / 256 < 2048 ? (n_pixels + 255) / 256 : 2048);
    if (dtype == TIF_DTYPE_F32)
        k_flow_rad_max<float><<<blocks, 256, 0, stream>>>((const float*)flow, n_pixels, use_clip, clip_lo, clip_hi, (unsigned long long*)out);
    else
        k_flow_rad_max<double><<<blocks, 256, 0, stream>>>((const double*)flow, n_pixels, use_clip, clip_lo, clip_hi, (unsigned long long*)out);
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

extern "C" int tif_flow_to_color_u8(const void* flow, int dtype, int64_t n_pixels, int use_clip, double clip_lo, double clip_hi,
                                    double norm, const double* wheel, int convert_to_bgr, uint8_t* out, void* stream_v) {
    if (!flow || !out || !wheel || n_pixels <= 0 || (dtype != TIF_DTYPE_F32 && dtype != TIF_DTYPE_F64)) {
        tif_set_error("tif_flow_to_color_u8: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    const unsigned blocks = (unsigned)((n_pixels + 255) / 256);
    if (dtype == TIF_DTYPE_F32)
        k_flow_to_color<float><<<blocks, 256, 0, stream>>>((const float*)flow, n_pixels, use_clip, clip_lo, clip_hi, norm, wheel, convert_to_bgr, out);
    else
        k_flow_to_color<double><<<blocks, 256, 0, stream>>>((const double*)flow, n_pixels, use_clip, clip_lo, clip_hi, norm, wheel, convert_to_bgr, out);
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}
