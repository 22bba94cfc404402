// tif_metrics.cu -- flow evaluation metrics, float64.
//
//   k_flow_error_mats   flow_evaluate.{available_pixel, EPE_mat, RMSE_mat, AAE_mat}   flow_evaluate.py:23-217
//   k_reduce_f64        the sums behind EPE / RMSE / AAE, deterministic two-level reduction
#include "tif_common.cuh"

#define TIF_UNKNOWN_FLOW_THRESH 1e7

__device__ __forceinline__ void endpoint_sph(double c, double r, double u, double v, int h, int w, double& theta,
                                             double& phi) {
    // flow_warp.flow_warp_meshgrid (flow_warp.py:37-49) then spherical_coordinates.erp2sph (:92-100)
    double eu = c + u, ev = r + v;
    if (eu >= (double)w - 0.5) eu -= (double)w;
    if (eu < -0.5) eu += (double)w;
    if (ev >= (double)h - 0.5) ev -= (double)h;
    if (ev < -0.5) ev += (double)h;
    theta = eu * (2 * TIF_PI / w) + TIF_PI / w - TIF_PI;
    phi = -(ev * (TIF_PI / h) + TIF_PI / h * 0.5) + 0.5 * TIF_PI;
    if (theta == TIF_PI) theta = -TIF_PI;
    if (phi == -0.5 * TIF_PI) phi = 0.5 * TIF_PI;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_flow_error_mats(const T* __restrict__ gt, const T* __restrict__ eva, const uint8_t* __restrict__ mask, int h, int w,
                  int spherical, double* __restrict__ epe, double* __restrict__ aae, uint8_t* __restrict__ valid) {
    const int64_t n_pix = (int64_t)h * w;
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (pix >= n_pix) return;
    const size_t gi = (size_t)b * n_pix + pix;
    double gu = (double)gt[2 * gi], gv = (double)gt[2 * gi + 1];
    double eu = (double)eva[2 * gi], ev = (double)eva[2 * gi + 1];
    // available_pixel (flow_evaluate.py:23-54), including its operator precedence: (not |u| > T) | (|v| > T)
    const bool unlarge = !(fabs(gu) > TIF_UNKNOWN_FLOW_THRESH) || (fabs(gv) > TIF_UNKNOWN_FLOW_THRESH);
    const bool unnan = !(isnan(gu) || isnan(gv));
    const bool unzero = (fabs(gu) > 0.0) || (fabs(gv) > 0.0);
    bool ok = unlarge && unnan && unzero;
    if (mask) ok = ok && (mask[gi] != 0);
    if (!ok) gu = gv = eu = ev = 0.0;                // the reference zeroes both flows where the ground truth is unusable
    valid[gi] = ok ? 1 : 0;
    double err;
    if (spherical) {
        const int r = (int)(pix / w), c = (int)(pix - (int64_t)r * w);
        double t1, p1, t2, p2;
        endpoint_sph((double)c, (double)r, gu, gv, h, w, t1, p1);
        endpoint_sph((double)c, (double)r, eu, ev, h, w, t2, p2);
        // great_circle_distance_uv (spherical_coordinates.py:45-52)
        const double s1 = sin((p2 - p1) * 0.5), s2 = sin((t2 - t1) * 0.5);
        const double a = s1 * s1 + cos(p1) * cos(p2) * (s2 * s2);
        err = fabs(1 * (2 * atan2(sqrt(a), sqrt(1 - a))));
    } else {
        const double du = gu - eu, dv = gv - ev;
        err = sqrt(du * du + dv * dv);
    }
    epe[gi] = err;
    if (aae) {                                       // AAE_mat planar branch (:209-214)
        const double cross = gu * ev - gv * eu, dot = gu * eu + gv * ev;
        aae[gi] = fabs(atan2(cross, dot));
    }
}

extern "C" int tif_flow_error_mats(const void* gt, const void* eva, int dtype, const uint8_t* mask, int batch, int h,
                                   int w, int spherical, double* epe, double* aae, uint8_t* valid, void* stream_v) {
    if (!gt || !eva || !epe || !valid || batch <= 0 || h <= 0 || w <= 0) {
        tif_set_error("tif_flow_error_mats: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    dim3 grid((unsigned)(((int64_t)h * w + 255) / 256), (unsigned)batch);
    if (dtype == TIF_DTYPE_F32)
        k_flow_error_mats<float><<<grid, 256, 0, stream>>>((const float*)gt, (const float*)eva, mask, h, w, spherical, epe, aae, valid);
    else if (dtype == TIF_DTYPE_F64)
        k_flow_error_mats<double><<<grid, 256, 0, stream>>>((const double*)gt, (const double*)eva, mask, h, w, spherical, epe, aae, valid);
    else
        return TIF_ERR_INVALID_ARG;
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

// sum (power 1) or sum of squares (power 2) of n doubles per image; partial sums per block, then a fixed-order pass
__global__ void __launch_bounds__(256)
k_reduce_partial(const double* __restrict__ x, int64_t n, int power, double* __restrict__ partial) {
    const int b = blockIdx.y;
    const double* xb = x + (size_t)b * n;
    double v = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double t = xb[i];
        v += (power == 2) ? t * t : t;
    }
    __shared__ double sh[8];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += sh[k];
        partial[(size_t)b * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void k_reduce_final(const double* __restrict__ partial, int n_blocks, double* __restrict__ out) {
    const int b = blockIdx.x, lane = threadIdx.x;
    double v = 0.0;
    for (int i = lane; i < n_blocks; i += 32) v += partial[(size_t)b * n_blocks + i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) out[b] = v;
}

#define TIF_REDUCE_BLOCKS 1184      // 8 x 148 SMs

extern "C" int64_t tif_reduce_f64_scratch_bytes(int batch) {
    return batch <= 0 ? 0 : (int64_t)batch * TIF_REDUCE_BLOCKS * (int64_t)sizeof(double);
}

extern "C" int tif_reduce_f64(const double* x, int batch, int64_t n, int power, double* out, void* scratch,
                              void* stream_v) {
    if (!x || !out || !scratch || batch <= 0 || n <= 0 || (power != 1 && power != 2)) {
        tif_set_error("tif_reduce_f64: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    k_reduce_partial<<<dim3(TIF_REDUCE_BLOCKS, (unsigned)batch), 256, 0, stream>>>(x, n, power, (double*)scratch);
    TIF_CUDA_CHECK(cudaGetLastError());
    k_reduce_final<<<batch, 32, 0, stream>>>((const double*)scratch, TIF_REDUCE_BLOCKS, out);
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

// ------------------------------------------------------------------------------------------------
// projection.get_blend_weight_ico / get_blend_weight_cubemap ("image_warp_error") -- projection.py:15-117
// ------------------------------------------------------------------------------------------------
// Stand-alone form of the photometric face weights (the stitch kernels fuse them): N source positions, N target
// positions, two [H,W,C] images.  scipy map_coordinates(order=1, mode='constant', cval=255) per channel in
// float64 with scipy's weight construction and accumulation order; a uint8 image rounds every sample half up
// (scipy's output dtype is the input's), a float32 one rounds to float32.
template <typename T>
__device__ __forceinline__ double sample_constant255(const T* __restrict__ img, int ch, int channels, int h, int w, double y,
                                                     double x) {
    if (!(y >= 0.0 && y <= (double)(h - 1) && x >= 0.0 && x <= (double)(w - 1))) return 255.0;
    const double y0f = floor(y), x0f = floor(x);
    const int y0 = (int)y0f, x0 = (int)x0f;
    const int y1 = y0 + 1 >= h ? (h - 2 > 0 ? h - 2 : 0) : y0 + 1;      // only reached with weight 0 (coordinate == n - 1)
    const int x1 = x0 + 1 >= w ? (w - 2 > 0 ? w - 2 : 0) : x0 + 1;
    const double fy = __dsub_rn(y, y0f), fx = __dsub_rn(x, x0f);
    const double wy0 = __dsub_rn(1.0, fy), wy1 = __dsub_rn(1.0, wy0), wx0 = __dsub_rn(1.0, fx), wx1 = __dsub_rn(1.0, wx0);
    const double p00 = (double)img[((size_t)y0 * w + x0) * channels + ch], p01 = (double)img[((size_t)y0 * w + x1) * channels + ch];
    const double p10 = (double)img[((size_t)y1 * w + x0) * channels + ch], p11 = (double)img[((size_t)y1 * w + x1) * channels + ch];
    double t = __dmul_rn(__dmul_rn(p00, wy0), wx0);
    t = __dadd_rn(t, __dmul_rn(__dmul_rn(p01, wy0), wx1));
    t = __dadd_rn(t, __dmul_rn(__dmul_rn(p10, wy1), wx0));
    t = __dadd_rn(t, __dmul_rn(__dmul_rn(p11, wy1), wx1));
    if (sizeof(T) == 1) {               // uint8: round half up, clamp
        t = t > 0.0 ? t + 0.5 : 0.0;
        t = t > 255.0 ? 255.0 : t;
        return (double)(uint8_t)t;
    }
    if (sizeof(T) == 4) return (double)(float)t;
    return t;
}

// per point: ico -> mean over channels of |tar - src| (np.mean(axis=1)), cubemap -> its L2 norm; and the block's
// partial sum of all |tar - src| (the face-global mean of the ico weights), fixed order
template <typename T>
__global__ void __launch_bounds__(256)
k_warp_error_points(const T* __restrict__ src, const T* __restrict__ tar, int h, int w, int channels,
                    const double* __restrict__ x_src, const double* __restrict__ y_src, const double* __restrict__ x_tar,
                    const double* __restrict__ y_tar, int64_t n, int kind, double* __restrict__ err,
                    double* __restrict__ partial) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double sum = 0.0, sq = 0.0;
        for (int ch = 0; ch < channels; ++ch) {
            const double t = sample_constant255(tar, ch, channels, h, w, y_tar[i], x_tar[i]);
            const double s = sample_constant255(src, ch, channels, h, w, y_src[i], x_src[i]);
            const double d = fabs(__dsub_rn(t, s));
            sum = __dadd_rn(sum, d);
            sq = __dadd_rn(sq, __dmul_rn(d, d));
        }
        err[i] = kind == 0 ? sum / (double)channels : sqrt(sq);
        acc += sum;
    }
    __shared__ double sh[8];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int wp = 0; wp < 8; ++wp) v += sh[wp];
        partial[blockIdx.x] = v;
    }
}

// ico: w = exp(-err / mean_all), mean_all = sum of the partials / (n channels); cubemap: w = err ? 0.95 / err : 1
__global__ void __launch_bounds__(256)
k_warp_error_weights(const double* __restrict__ partial, int n_blocks, int64_t n, int channels, int kind,
                     double* __restrict__ err_to_weight) {
    __shared__ double mean_all;
    if (kind == 0) {
        if (threadIdx.x < 32) {
            double v = 0.0;
            for (int i = threadIdx.x; i < n_blocks; i += 32) v += partial[i];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (threadIdx.x == 0) mean_all = v / ((double)n * (double)channels);
        }
        __syncthreads();
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double e = err_to_weight[i];
        err_to_weight[i] = kind == 0 ? exp(-(e / mean_all)) : (e != 0.0 ? 0.95 / e : 1.0);
    }
}

static int warp_error_blocks(int64_t n) {
    const int64_t want = (n + 255) / 256;
    return (int)(want < 148 * 8 ? (want > 0 ? want : 1) : 148 * 8);
}

extern "C" int64_t tif_blend_weight_points_scratch_bytes(int64_t n) {
    return n > 0 ? (int64_t)warp_error_blocks(n) * (int64_t)sizeof(double) : 0;
}

extern "C" int tif_blend_weight_points(const void* src, const void* tar, int image_dtype, int h, int w, int channels,
                                       const double* x_src, const double* y_src, const double* x_tar, const double* y_tar,
                                       int64_t n, int weight_kind, double* weights, void* scratch, void* stream_v) {
    if (!src || !tar || !x_src || !y_src || !x_tar || !y_tar || !weights || !scratch || n <= 0 || h <= 0 || w <= 0 ||
        channels <= 0 || (weight_kind != 0 && weight_kind != 1)) {
        tif_set_error("tif_blend_weight_points: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int n_blocks = warp_error_blocks(n);
    double* partial = (double*)scratch;
#define WE_ARGS h, w, channels, x_src, y_src, x_tar, y_tar, n, weight_kind, weights, partial
    if (image_dtype == TIF_DTYPE_U8)
        k_warp_error_points<uint8_t><<<n_blocks, 256, 0, stream>>>((const uint8_t*)src, (const uint8_t*)tar, WE_ARGS);
    else if (image_dtype == TIF_DTYPE_F32)
        k_warp_error_points<float><<<n_blocks, 256, 0, stream>>>((const float*)src, (const float*)tar, WE_ARGS);
    else if (image_dtype == TIF_DTYPE_F64)
        k_warp_error_points<double><<<n_blocks, 256, 0, stream>>>((const double*)src, (const double*)tar, WE_ARGS);
    else {
        tif_set_error("tif_blend_weight_points: unknown image dtype %d", image_dtype);
        return TIF_ERR_INVALID_ARG;
    }
#undef WE_ARGS
    TIF_CUDA_CHECK(cudaGetLastError());
    k_warp_error_weights<<<n_blocks, 256, 0, stream>>>(partial, n_blocks, n, channels, weight_kind, weights);
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}
