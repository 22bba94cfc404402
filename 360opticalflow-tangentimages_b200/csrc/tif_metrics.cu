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
