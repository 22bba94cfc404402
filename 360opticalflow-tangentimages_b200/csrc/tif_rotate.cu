// tif_rotate.cu -- global-rotation helpers of the pipeline, float64.
//
//   k_rotation_mv         spherical_coordinates.rotation2erp_motion_vector   spherical_coordinates.py:185-218
//   k_cross_cov (+reduce) the 3x3 cross covariance of flow_warp.flow2rotation_3d  flow_warp.py:103-127, pointcloud_utils.py:29
//   k_rotate_endpoint     projection.flow_rotate_endpoint                    projection.py:120-159
//
// rotate_erp_array (spherical_coordinates.py:172-182) = k_rotation_mv + the existing k_warp_backward with a float64
// flow, so the rotated uint8 image is rounded exactly like scipy's.  The 3x3 SVD stays on the host (9 numbers).
#include "tif_common.cuh"

__device__ __forceinline__ void erp2sph_px(double x, double y, int h, int w, double& theta, double& phi) {
    // spherical_coordinates.py:92-100
    theta = x * (2 * TIF_PI / w) + TIF_PI / w - TIF_PI;
    phi = -(y * (TIF_PI / h) + TIF_PI / h * 0.5) + 0.5 * TIF_PI;
    if (theta == TIF_PI) theta = -TIF_PI;
    if (phi == -0.5 * TIF_PI) phi = 0.5 * TIF_PI;
}

struct Rot3 {
    double m[9];
};

__global__ void __launch_bounds__(256)
k_rotation_mv(Rot3 R, int h, int w, int wraparound, double* __restrict__ out) {
    const int64_t n_pix = (int64_t)h * w;
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= n_pix) return;
    const int r = (int)(pix / w), c = (int)(pix - (int64_t)r * w);
    double theta, phi;
    erp2sph_px((double)c, (double)r, h, w, theta, phi);
    // sph2car (:151-169), rotate, car2sph (:128-148)
    double sp, cp, st, ct;
    sincos(phi, &sp, &cp);
    sincos(theta, &st, &ct);
    const double x = 1.0 * cp * st, y = -1.0 * sp, z = 1.0 * cp * ct;
    const double rx = R.m[0] * x + R.m[1] * y + R.m[2] * z;
    const double ry = R.m[3] * x + R.m[4] * y + R.m[5] * z;
    const double rz = R.m[6] * x + R.m[7] * y + R.m[8] * z;
    const double radius = sqrt(rx * rx + ry * ry + rz * rz);
    double t2 = 0.0, p2 = 0.0;
    if (radius > 1e-10) {
        t2 = atan2(rx, rz);
        p2 = -asin(ry / radius);
    }
    // sph2erp without modulo (:122-124)
    const double ex = (t2 + TIF_PI) / (2.0 * TIF_PI / w) - 0.5;
    const double ey = (-p2 + 0.5 * TIF_PI) / (TIF_PI / h) - 0.5;
    double mx = ex - (double)c;
    const double my = ey - (double)r;
    if (wraparound) {       // sph_wraparound, :242-258
        const bool long_line = fabs(t2 - theta) > TIF_PI;
        if (long_line && theta < 0 && t2 > 0) mx -= w;
        else if (long_line && theta > 0 && t2 < 0) mx += w;
    }
    out[2 * pix] = mx;
    out[2 * pix + 1] = my;
}

extern "C" int tif_rotation_motion_vector_f64(const double* rotation, int h, int w, int wraparound, double* out,
                                              void* stream) {
    if (!rotation || !out || h <= 0 || w <= 0) {
        tif_set_error("tif_rotation_motion_vector_f64: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    Rot3 R;
    for (int i = 0; i < 9; ++i) R.m[i] = rotation[i];
    k_rotation_mv<<<(unsigned)(((int64_t)h * w + 255) / 256), 256, 0, (cudaStream_t)stream>>>(R, h, w, wraparound, out);
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

// ------------------------------------------------------------------------------------------------
// cross covariance: one block = 256 pixels of the middle rows; block-level tree reduction into partial[b][block][9],
// then one block per image adds the partials in a fixed order (deterministic, unlike float atomics)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_cross_cov(const T* __restrict__ flow, int h, int w, int row0, int row1, double* __restrict__ partial) {
    const int b = blockIdx.y;
    const int64_t n_sel = (int64_t)(row1 - row0) * w;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0;
    if (i < n_sel) {
        const int r = row0 + (int)(i / w), c = (int)(i % w);
        const size_t gi = ((size_t)b * h + r) * w + c;
        double ts, ps, tt, pt;
        erp2sph_px((double)c, (double)r, h, w, ts, ps);
        // flow_warp_meshgrid (flow_warp.py:37-49): one wrap each way
        double eu = (double)c + (double)flow[2 * gi], ev = (double)r + (double)flow[2 * gi + 1];
        if (eu >= (double)w - 0.5) eu -= (double)w;
        if (eu < -0.5) eu += (double)w;
        if (ev >= (double)h - 0.5) ev -= (double)h;
        if (ev < -0.5) ev += (double)h;
        erp2sph_px(eu, ev, h, w, tt, pt);
        double s3[3], t3[3], sn, cs;
        sincos(ps, &sn, &cs);
        s3[0] = 1.0 * cs * sin(ts);
        s3[1] = -1.0 * sn;
        s3[2] = 1.0 * cs * cos(ts);
        sincos(pt, &sn, &cs);
        t3[0] = 1.0 * cs * sin(tt);
        t3[1] = -1.0 * sn;
        t3[2] = 1.0 * cs * cos(tt);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int d = 0; d < 3; ++d) acc[a * 3 + d] = s3[a] * t3[d];
    }
    __shared__ double sh[8][9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double v = 0.0;
        for (int wp = 0; wp < 8; ++wp) v += sh[wp][threadIdx.x];
        partial[((size_t)b * gridDim.x + blockIdx.x) * 9 + threadIdx.x] = v;
    }
}

__global__ void k_cross_cov_reduce(const double* __restrict__ partial, int n_blocks, double* __restrict__ out) {
    // one block of 9 x 32 threads per image: thread (k, lane) sums a strided slice, then a fixed-order tree
    const int b = blockIdx.x, k = threadIdx.x / 32, lane = threadIdx.x & 31;
    double v = 0.0;
    for (int i = lane; i < n_blocks; i += 32) v += partial[((size_t)b * n_blocks + i) * 9 + k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) out[b * 9 + k] = v;
}

extern "C" int64_t tif_flow_cross_covariance_scratch_bytes(int batch, int h, int w) {
    if (batch <= 0 || h <= 0 || w <= 0) return 0;
    const int row0 = (int)(h * 0.25), row1 = (int)(h * 0.75);
    const int64_t n_blocks = ((int64_t)(row1 - row0) * w + 255) / 256;
    return (int64_t)batch * n_blocks * 9 * (int64_t)sizeof(double);
}

extern "C" int tif_flow_cross_covariance(const void* flow, int dtype, int batch, int h, int w, double* out,
                                         void* scratch, void* stream_v) {
    if (!flow || !out || !scratch || batch <= 0 || h < 4 || w <= 0) {
        tif_set_error("tif_flow_cross_covariance: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int row0 = (int)(h * 0.25), row1 = (int)(h * 0.75);      // flow_warp.py:117-118
    const int n_blocks = (int)(((int64_t)(row1 - row0) * w + 255) / 256);
    dim3 grid((unsigned)n_blocks, (unsigned)batch);
    if (dtype == TIF_DTYPE_F32)
        k_cross_cov<float><<<grid, 256, 0, stream>>>((const float*)flow, h, w, row0, row1, (double*)scratch);
    else if (dtype == TIF_DTYPE_F64)
        k_cross_cov<double><<<grid, 256, 0, stream>>>((const double*)flow, h, w, row0, row1, (double*)scratch);
    else
        return TIF_ERR_INVALID_ARG;
    TIF_CUDA_CHECK(cudaGetLastError());
    k_cross_cov_reduce<<<batch, 9 * 32, 0, stream>>>((const double*)scratch, n_blocks, out);
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

// ------------------------------------------------------------------------------------------------
// flow2rotation_2d statistics -- flow_warp.py:136-187 (use_weight=False, the only form that runs in the reference)
// ------------------------------------------------------------------------------------------------
// Six sums per image in one pass: [0] sum of u over the whole image (theta_delta = mean(2 pi u / W), :152-153), and
// over the columns [col0, col1) of the centre band (:157-170): [1] sum of sign(v), [2] sum and [3] count of v > 0,
// [4] sum and [5] count of v < 0.  The band depends on [0], so the caller runs the pass twice.  Fixed-order
// two-level float64 reduction: bitwise reproducible.
template <typename T>
__global__ void __launch_bounds__(256) k_rot2d_stats(const T* __restrict__ flow, int h, int w, int col0, int col1,
                                                     double* __restrict__ partial) {
    const int b = blockIdx.y;
    const int64_t n = (int64_t)h * w;
    const T* f = flow + (size_t)b * n * 2;
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % w);
        const double u = (double)f[2 * i], v = (double)f[2 * i + 1];
        acc[0] += u;
        if (c >= col0 && c < col1) {
            if (v > 0.0) {
                acc[1] += 1.0;
                acc[2] += v;
                acc[3] += 1.0;
            } else if (v < 0.0) {
                acc[1] -= 1.0;
                acc[4] += v;
                acc[5] += 1.0;
            }
        }
    }
    __shared__ double sh[8][6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = 0.0;
        for (int wp = 0; wp < 8; ++wp) v += sh[wp][threadIdx.x];
        partial[((size_t)b * gridDim.x + blockIdx.x) * 6 + threadIdx.x] = v;
    }
}

__global__ void k_rot2d_reduce(const double* __restrict__ partial, int n_blocks, double* __restrict__ out) {
    const int b = blockIdx.x, k = threadIdx.x / 32, lane = threadIdx.x & 31;
    double v = 0.0;
    for (int i = lane; i < n_blocks; i += 32) v += partial[((size_t)b * n_blocks + i) * 6 + k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) out[b * 6 + k] = v;
}

static int rot2d_blocks(int h, int w) {
    const int64_t want = ((int64_t)h * w + 255) / 256;
    return (int)(want < 148 * 8 ? want : 148 * 8);
}

extern "C" int64_t tif_flow_rotation2d_stats_scratch_bytes(int batch, int h, int w) {
    if (batch <= 0 || h <= 0 || w <= 0) return 0;
    return (int64_t)batch * rot2d_blocks(h, w) * 6 * (int64_t)sizeof(double);
}

extern "C" int tif_flow_rotation2d_stats(const void* flow, int dtype, int batch, int h, int w, int col_start, int col_stop,
                                         double* out, void* scratch, void* stream_v) {
    if (!flow || !out || !scratch || batch <= 0 || h <= 0 || w <= 0) {
        tif_set_error("tif_flow_rotation2d_stats: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int n_blocks = rot2d_blocks(h, w);
    dim3 grid((unsigned)n_blocks, (unsigned)batch);
    if (dtype == TIF_DTYPE_F32)
        k_rot2d_stats<float><<<grid, 256, 0, stream>>>((const float*)flow, h, w, col_start, col_stop, (double*)scratch);
    else if (dtype == TIF_DTYPE_F64)
        k_rot2d_stats<double><<<grid, 256, 0, stream>>>((const double*)flow, h, w, col_start, col_stop, (double*)scratch);
    else
        return TIF_ERR_INVALID_ARG;
    TIF_CUDA_CHECK(cudaGetLastError());
    k_rot2d_reduce<<<batch, 6 * 32, 0, stream>>>((const double*)scratch, n_blocks, out);
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

// ------------------------------------------------------------------------------------------------
// flow_rotate_endpoint
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double bilinear_wrap_f64(const double* __restrict__ field, int ch, double y, double x, int h,
                                                    int w) {
    // scipy map_coordinates(order=1, mode='wrap'), float64 in and out, scipy's accumulation order
    x = scipy_wrap(x, w);
    y = scipy_wrap(y, h);
    double x0 = floor(x), y0 = floor(y);
    if (x0 > (double)(w - 2)) x0 = (double)(w - 2);
    if (y0 > (double)(h - 2)) y0 = (double)(h - 2);
    if (x0 < 0.0) x0 = 0.0;
    if (y0 < 0.0) y0 = 0.0;
    const double fx = x - x0, fy = y - y0;
    const double wx0 = __dsub_rn(1.0, fx), wx1 = __dsub_rn(1.0, wx0), wy0 = __dsub_rn(1.0, fy), wy1 = __dsub_rn(1.0, wy0);
    const double* p = field + (((size_t)y0 * w + (size_t)x0) * 2 + ch);
    const size_t row = (size_t)w * 2;
    double t = __dmul_rn(__dmul_rn(p[0], wy0), wx0);
    t = __dadd_rn(t, __dmul_rn(__dmul_rn(p[2], wy0), wx1));
    t = __dadd_rn(t, __dmul_rn(__dmul_rn(p[row], wy1), wx0));
    t = __dadd_rn(t, __dmul_rn(__dmul_rn(p[row + 2], wy1), wx1));
    return t;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_rotate_endpoint(const T* __restrict__ flow, const double* __restrict__ field, T* __restrict__ out, int h, int w,
                  int wraparound) {
    const int64_t n_pix = (int64_t)h * w;
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (pix >= n_pix) return;
    const int r = (int)(pix / w), c = (int)(pix - (int64_t)r * w);
    const size_t gi = (size_t)b * n_pix + pix;
    // end point in the image (flow_postproc.erp_pixles_modulo, projection.py:137)
    double ex = np_remainder((double)c + (double)flow[2 * gi] + 0.5, (double)w) - 0.5;
    double ey = np_remainder((double)r + (double)flow[2 * gi + 1] + 0.5, (double)h) - 0.5;
    const double ru = bilinear_wrap_f64(field, 0, ey, ex, h, w);
    const double rv = bilinear_wrap_f64(field, 1, ey, ex, h, w);
    ex = np_remainder(ex + ru + 0.5, (double)w) - 0.5;
    ey = np_remainder(ey + rv + 0.5, (double)h) - 0.5;
    double u = ex - (double)c;
    const double v = ey - (double)r;
    if (wraparound) {       // flow_postproc.erp_of_wraparound, default threshold W/2
        const double thr = (double)w / 2.0;
        const int split = (int)((double)w - 1.0 - thr);
        if (c < split) {
            if (u > thr) u -= (double)w;
        } else if (c < w - 1) {
            if (u < -thr) u += (double)w;
        }
    }
    out[2 * gi] = (T)u;
    out[2 * gi + 1] = (T)v;
}

extern "C" int tif_flow_rotate_endpoint(const void* flow, int dtype, const double* rotation_field, int batch, int h,
                                        int w, int wraparound, void* out, void* stream_v) {
    if (!flow || !rotation_field || !out || batch <= 0 || h < 2 || w < 2) {
        tif_set_error("tif_flow_rotate_endpoint: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    dim3 grid((unsigned)(((int64_t)h * w + 255) / 256), (unsigned)batch);
    if (dtype == TIF_DTYPE_F32)
        k_rotate_endpoint<float><<<grid, 256, 0, stream>>>((const float*)flow, rotation_field, (float*)out, h, w, wraparound);
    else if (dtype == TIF_DTYPE_F64)
        k_rotate_endpoint<double><<<grid, 256, 0, stream>>>((const double*)flow, rotation_field, (double*)out, h, w, wraparound);
    else
        return TIF_ERR_INVALID_ARG;
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}
