// tif_plan.cu -- geometry plan kernels (fp64) and plan lifetime.  COMPILE WITH -fmad=false.
//
// The plan is the input-independent half of the reference's two hot functions, which the reference
// recomputes for every face of every call:
//   stage 1  erp2ico_image   projection_icosahedron.py:197-229  tangent pixel -> ERP sample position
//   stage 2  ico2erp_flow    projection_icosahedron.py:282-329  ERP pixel -> accepting faces + nearest tangent pixel
// Both are evaluated once per configuration by the kernels below and kept resident in HBM.
//
// Bit-exactness: face membership and pixel indices are decided by float64 compares.  Every
// transcendental those compares depend on comes from host tables evaluated by numpy (TifPlanTables),
// and this file is compiled without FMA contraction, so each + - * / below rounds once, in the
// reference's operation order.  Stage 1 needs sqrt/atan2/sin/cos/asin of 2-D quantities; those use
// CUDA's fp64 libm (<= 2 ulp from numpy's), which moves a sample position by ~1e-13 px.
#include <float.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "tif_common.cuh"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_last_error[512] = "";

void tif_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

extern "C" int tif_abi_version(void) { return TIF_ABI_VERSION; }

extern "C" const char* tif_last_error(void) { return g_last_error; }

extern "C" int tif_struct_sizes(int32_t* face_bytes, int32_t* tables_bytes, int32_t* info_bytes) {
    if (face_bytes) *face_bytes = (int32_t)sizeof(TifFace);
    if (tables_bytes) *tables_bytes = (int32_t)sizeof(TifPlanTables);
    if (info_bytes) *info_bytes = (int32_t)sizeof(TifPlanInfo);
    return TIF_OK;
}

extern "C" const char* tif_status_string(int status) {
    switch (status) {
        case TIF_OK: return "TIF_OK";
        case TIF_ERR_INVALID_ARG: return "TIF_ERR_INVALID_ARG";
        case TIF_ERR_CUDA: return "TIF_ERR_CUDA";
        case TIF_ERR_ALLOC: return "TIF_ERR_ALLOC";
        case TIF_ERR_RANGE: return "TIF_ERR_RANGE";
        case TIF_ERR_UNSUPPORTED: return "TIF_ERR_UNSUPPORTED";
        default: return "TIF_ERR_UNKNOWN";
    }
}

// ------------------------------------------------------------------------------------------------
// point-in-polygon, polygon.py:59-118 with on_line=True (ray cast, inclusive y range, eps-wide edges)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pip_accept(double px, double py, const double (*tri)[2], int n, double eps) {
    bool inside = false, online = false;
    for (int k = 0; k < n; ++k) {
        const int k2 = (k + 1 == n) ? 0 : k + 1;
        const double x1 = tri[k][0], y1 = tri[k][1];
        const double x2 = tri[k2][0], y2 = tri[k2][1];
        bool t = (py >= fmin(y1, y2)) && (py <= fmax(y1, y2)) && (px <= fmax(x1, x2));
        double ix;
        if (fabs(y1 - y2) <= eps) {
            t = t && (px >= fmin(x1, x2));
            ix = px;
        } else {
            ix = (py - y1) * (x2 - x1) / (y2 - y1) + x1;
        }
        online = online || (t && (fabs(px - ix) <= eps));
        inside = inside != (t && (px <= ix));
    }
    return inside || online;
}

// ------------------------------------------------------------------------------------------------
// stage 1: one thread per tangent pixel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_plan_stage1(const TifFace* __restrict__ faces, const int32_t* __restrict__ row_face,
              const double* __restrict__ gnom_x, const double* __restrict__ gnom_y,
              int erp_h, int erp_w, int tangent_w, int64_t n_pix, int kind,
              uint32_t* __restrict__ s1_base, double2* __restrict__ s1_frac64,
              double* __restrict__ s1_ex, double* __restrict__ s1_ey) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pix) return;
    const int64_t grow = p / tangent_w;          // row in the concatenated stack
    const int j = (int)(p - grow * tangent_w);
    const int f = row_face[grow];
    const TifFace& F = faces[f];
    const double x = gnom_x[(int64_t)f * tangent_w + j];
    const double y = gnom_y[grow];

    // reverse_gnomonic_projection, gnomonic_projection.py:76-93
    double rho = sqrt(x * x + y * y);
    const bool zero = (rho == 0.0);
    if (zero) rho = DBL_EPSILON;
    const double c = atan2(rho, 1.0);
    double sc, cc;
    sincos(c, &sc, &cc);
    double phi = asin(cc * F.sin_phi0 + (y * sc * F.cos_phi0) / rho);
    double theta = F.theta0 + atan2(x * sc, rho * F.cos_phi0 * cc - y * F.sin_phi0 * sc);
    if (zero) {
        phi = F.phi0;
        theta = F.theta0;
    }
    const double theta_raw = theta, phi_raw = phi;
    // sph2erp(sph_modulo=True), spherical_coordinates.py:59-60,122-124
    theta = np_remainder(theta + TIF_PI, 2 * TIF_PI) - TIF_PI;
    phi = -(np_remainder(-phi + 0.5 * TIF_PI, TIF_PI) - 0.5 * TIF_PI);
    const double ex = (theta + TIF_PI) / (2.0 * TIF_PI / erp_w) - 0.5;
    const double ey = (-phi + 0.5 * TIF_PI) / (TIF_PI / erp_h) - 0.5;
    s1_ex[p] = ex;
    s1_ey[p] = ey;

    // resampling position: the icosahedron stage samples at the pixel-centre coordinates above; the cubemap stage
    // (projection_cubemap.py:169-178) uses x = (theta+pi)/(2 pi) W, y = (pi/2-phi)/pi H of the continuous theta
    // with a single fold into the image
    double rx = ex, ry = ey;
    if (kind == TIF_PLAN_CUBEMAP) {
        rx = ((theta_raw + TIF_PI) / (2 * TIF_PI)) * erp_w;
        ry = (-phi_raw + TIF_PI / 2.0) / TIF_PI * erp_h;
        if (rx < 0) rx = rx + erp_w;
        if (rx >= erp_w) rx = rx - erp_w;
        if (ry < 0) ry = ry + erp_h;
        if (ry >= erp_h) ry = ry - erp_h;
    }
    // scipy map_coordinates(order=1, mode='wrap'): period n-1, taps floor(c), floor(c)+1.
    // The base tap is clamped to n-2 (fraction 1.0 when c == n-1 exactly) so both taps are always
    // in range; the product with scipy's zero weight is unchanged.
    const double cx = scipy_wrap(rx, erp_w);
    const double cy = scipy_wrap(ry, erp_h);
    double x0 = floor(cx), y0 = floor(cy);
    if (x0 > (double)(erp_w - 2)) x0 = (double)(erp_w - 2);
    if (y0 > (double)(erp_h - 2)) y0 = (double)(erp_h - 2);
    if (x0 < 0.0) x0 = 0.0;
    if (y0 < 0.0) y0 = 0.0;
    const double fx = cx - x0, fy = cy - y0;
    uint32_t base = (uint32_t)((int64_t)y0 * erp_w + (int64_t)x0);
    if (!pip_accept(x, y, F.tri, F.n_vertices, F.eps_image)) base |= TIF_S1_OUTSIDE;
    s1_base[p] = base;
    s1_frac64[p] = make_double2(fx, fy);
}

// ------------------------------------------------------------------------------------------------
// stage 2: one thread per ERP pixel, faces visited in index order ("pull" form)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int floor_div(int a, int b) {   // b > 0
    int q = a / b, r = a % b;
    return (r != 0 && r < 0) ? q - 1 : q;
}

// number of integers k in [start, stop] with k mod n == v  (0 <= v < n): multiplicity of a wrapped
// candidate row / column after flow_postproc.erp_pixles_modulo (flow_postproc.py:12-13)
__device__ __forceinline__ int wrapped_multiplicity(int start, int stop, int v, int n) {
    if (stop < start) return 0;
    return floor_div(stop - v, n) - floor_div(start - 1 - v, n);
}

template <bool FILL>
__global__ void __launch_bounds__(256)
k_plan_stage2(const TifFace* __restrict__ faces, int n_faces, int erp_h, int erp_w, int tangent_w, int kind,
              const double* __restrict__ row_sincos, const double* __restrict__ col_sincos,
              uint8_t* __restrict__ count, uint32_t* __restrict__ entries, int max_k,
              uint8_t* __restrict__ ref_mark, unsigned long long* __restrict__ face_mult_sum,
              const double* __restrict__ s1_ex, const double* __restrict__ s1_ey, float2* __restrict__ snap,
              int* __restrict__ status /* [0]=max count, [1]=range errors */) {
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_pix = (int64_t)erp_h * erp_w;
    if (pix >= n_pix) return;
    const int r = (int)(pix / erp_w);
    const int c = (int)(pix - (int64_t)r * erp_w);
    const double sin_phi = row_sincos[2 * r], cos_phi = row_sincos[2 * r + 1];
    int n = 0;
    for (int f = 0; f < n_faces; ++f) {
        const TifFace& F = faces[f];
        const int mr = wrapped_multiplicity(F.row_start, F.row_stop, r, erp_h);
        if (mr == 0) continue;
        const int mc = wrapped_multiplicity(F.col_start, F.col_stop, c, erp_w);
        if (mc == 0) continue;
        const double sin_d = col_sincos[2 * ((int64_t)f * erp_w + c)];
        const double cos_d = col_sincos[2 * ((int64_t)f * erp_w + c) + 1];
        // gnomonic_projection, gnomonic_projection.py:36-49 (hemisphere mask discarded by the caller)
        double cos_c = F.sin_phi0 * sin_phi + F.cos_phi0 * cos_phi * cos_d;
        double x, y;
        if (cos_c == 0.0) {
            x = 0.0;
            y = 0.0;
        } else {
            x = cos_phi * sin_d / cos_c;
            y = (F.cos_phi0 * sin_phi - F.sin_phi0 * cos_phi * cos_d) / cos_c;
        }
        if (!pip_accept(x, y, F.tri, F.n_vertices, F.eps_stitch)) continue;
        if (FILL) {
            // gnomonic2pixel, gnomonic_projection.py:141-147 (padding_size 0.0, padded range)
            const double fx = (x - F.xmin + 0.0) * F.g2p_x + 0.5;
            const double fy = -(y - F.ymax - 0.0) * F.g2p_y + 0.5;
            int ix = (int)fx, iy = (int)fy;          // astype(int): truncation toward zero
            int m = mr * mc;
            if (kind == TIF_PLAN_CUBEMAP)       // no face-global mean there: the field carries the inner-square flag
                m = (F.n_inner > 0 && pip_accept(x, y, F.inner, F.n_inner, F.eps_inner)) ? 1 : 0;
            if (ix < 0 || ix >= tangent_w || iy < 0 || iy >= F.tangent_h || m > 7) {
                atomicAdd(&status[1], 1);
                ix = min(max(ix, 0), tangent_w - 1);
                iy = min(max(iy, 0), F.tangent_h - 1);
            }
            const int64_t tpix = (int64_t)F.pix_offset + (int64_t)iy * tangent_w + ix;
            if (n < max_k) {
                entries[(int64_t)n * n_pix + pix] =
                    ((uint32_t)f << 25) | ((uint32_t)min(m, 7) << 22) | ((uint32_t)iy << 11) | (uint32_t)ix;
                // where the nearest tangent pixel itself sits in the ERP image (stage 1 of this plan), relative
                // to this ERP pixel: the constant part of every flow that is gathered through this entry
                double sx = s1_ex[tpix] - (double)c;
                if (sx > 0.5 * erp_w) sx -= (double)erp_w;
                else if (sx < -0.5 * erp_w) sx += (double)erp_w;
                snap[(int64_t)n * n_pix + pix] = make_float2((float)sx, (float)(s1_ey[tpix] - (double)r));
            }
            ref_mark[tpix] = 1;
            atomicAdd(&face_mult_sum[f], (unsigned long long)m);
        }
        ++n;
    }
    if (FILL) {
        count[pix] = (uint8_t)min(n, max_k);
    } else {
        if (n > 0) atomicMax(&status[0], n);
    }
}

__global__ void k_count_marks(const uint8_t* __restrict__ mark, int64_t n, unsigned long long* __restrict__ out) {
    unsigned long long local = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        local += mark[i];
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

// ------------------------------------------------------------------------------------------------
// normwarp sample list: the accepted (ERP pixel, face) samples regrouped by the tangent pixel they snap to, in the
// order of the lift kernel's work list.  k_lift_nw lifts TIF_NW_BLOCK tangent pixels per block and then walks the
// samples of exactly those pixels (projection.py:52-58 needs |tar(end point) - src(ERP pixel)| per sample).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sample_ref_index(uint32_t e, const TifFace* __restrict__ faces, int tangent_w,
                                                const int32_t* __restrict__ ref_index) {
    const int64_t tpix = (int64_t)faces[TIF_E_FACE(e)].pix_offset + (int64_t)TIF_E_IY(e) * tangent_w + TIF_E_IX(e);
    return ref_index[tpix];
}

template <bool FILL>
__global__ void __launch_bounds__(256)
k_plan_samples(const uint8_t* __restrict__ count, const uint32_t* __restrict__ entries, const TifFace* __restrict__ faces,
               const int32_t* __restrict__ ref_index, int erp_h, int erp_w, int tangent_w, uint32_t* __restrict__ per_ref,
               const uint32_t* __restrict__ start, uint2* __restrict__ smp) {
    const int64_t n_pix = (int64_t)erp_h * erp_w;
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= n_pix) return;
    const int r = (int)(pix / erp_w), c = (int)(pix - (int64_t)r * erp_w);
    const int n = count[pix];
    for (int k = 0; k < n; ++k) {
        const uint32_t e = entries[(int64_t)k * n_pix + pix];
        const int ref = sample_ref_index(e, faces, tangent_w, ref_index);
        const uint32_t slot = atomicAdd(&per_ref[ref], 1u);
        if (FILL)
            smp[start[ref] + slot] = make_uint2(((uint32_t)r << 16) | (uint32_t)c,
                                                (TIF_E_FACE(e) << 24) | ((uint32_t)k << 16) | (TIF_E_MULT(e) << 8) | ((uint32_t)ref % TIF_NW_BLOCK));
    }
}

// own ERP position of every work-list entry, split into integer and fraction (the lift adds the displacement to it
// when it samples the target image; float64 only here, once per plan)
__global__ void __launch_bounds__(256)
k_plan_ref_pos(const int32_t* __restrict__ ref_list, int n_ref, const LiftFace* __restrict__ lift_faces, int tangent_w,
               const double* __restrict__ s1_ex, const double* __restrict__ s1_ey, int4* __restrict__ ref_pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ref) return;
    const uint32_t ref = (uint32_t)ref_list[i];
    const int64_t p = (int64_t)(lift_faces[TIF_E_FACE(ref)].first_row + (int)TIF_E_IY(ref)) * tangent_w + (int)TIF_E_IX(ref);
    const double ex = s1_ex[p], ey = s1_ey[p];
    const double fx = floor(ex), fy = floor(ey);
    // x, y in [-1, 65534]: one word, biased by 1
    ref_pos[i] = make_int4(((int)fx + 1) | (((int)fy + 1) << 16), __float_as_int((float)(ex - fx)), __float_as_int((float)(ey - fy)), (int)p);
}

// ------------------------------------------------------------------------------------------------
// plan lifetime
// ------------------------------------------------------------------------------------------------
template <typename T>
static int dev_alloc(T** p, size_t n, TifPlan* plan) {
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
    if (e != cudaSuccess) {
        tif_set_error("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
        *p = nullptr;
        return TIF_ERR_ALLOC;
    }
    plan->device_bytes += (int64_t)(n * sizeof(T));
    return TIF_OK;
}

extern "C" int tif_plan_destroy(TifPlan* plan) {
    if (!plan) return TIF_OK;
    cudaFree(plan->d_faces);
    cudaFree(plan->d_lift_faces);
    cudaFree(plan->d_pix_face);
    cudaFree(plan->d_s1_base);
    cudaFree(plan->d_s1_frac64);
    cudaFree(plan->d_s1_ex);
    cudaFree(plan->d_s1_ey);
    cudaFree(plan->d_s1_tiles);
    cudaFree(plan->d_s2_count);
    cudaFree(plan->d_s2_entries);
    cudaFree(plan->d_s2_snap);
    cudaFree(plan->d_ref_mark);
    cudaFree(plan->d_ref_list);
    cudaFree(plan->d_face_mult_sum);
    cudaFree(plan->d_ref_pos);
    cudaFree(plan->d_smp);
    cudaFree(plan->d_blk_smp);
    free(plan->h_faces);
    free(plan);
    return TIF_OK;
}

// device scratch of tif_plan_create, released on every exit path
struct PlanScratch {
    double *gx = nullptr, *gy = nullptr, *rows = nullptr, *cols = nullptr;
    int* status = nullptr;
    unsigned long long* msum = nullptr;
    int32_t* ref_index = nullptr;
    uint32_t *per_ref = nullptr, *smp_start = nullptr;
    ~PlanScratch() {
        cudaFree(gx);
        cudaFree(gy);
        cudaFree(rows);
        cudaFree(cols);
        cudaFree(status);
        cudaFree(msum);
        cudaFree(ref_index);
        cudaFree(per_ref);
        cudaFree(smp_start);
    }
};

#define PLAN_TRY(expr)              \
    do {                            \
        int _s = (expr);            \
        if (_s != TIF_OK) {         \
            tif_plan_destroy(plan); \
            return _s;              \
        }                           \
    } while (0)

static int cuda_status(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return TIF_OK;
    tif_set_error("%s failed: %s", what, cudaGetErrorString(e));
    return TIF_ERR_CUDA;
}

template <typename T>
static int scratch_alloc(T** p, size_t n) {
    return cuda_status(cudaMalloc((void**)p, n * sizeof(T)), "cudaMalloc (plan scratch)");
}

extern "C" int tif_plan_create(const TifFace* faces, int n_faces, int erp_h, int erp_w, int tangent_w,
                               const TifPlanTables* tables, int kind, void* stream_v, TifPlan** plan_out) {
    if (!faces || !tables || !plan_out || n_faces <= 0 || n_faces > TIF_MAX_FACES || erp_h < 2 || erp_w < 2 ||
        tangent_w < 2 || tangent_w > TIF_MAX_TANGENT_DIM || !tables->gnom_x || !tables->gnom_y ||
        !tables->row_sincos || !tables->col_sincos) {
        tif_set_error("tif_plan_create: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    if (kind != TIF_PLAN_ICOSAHEDRON && kind != TIF_PLAN_CUBEMAP) {
        tif_set_error("tif_plan_create: unknown plan kind %d", kind);
        return TIF_ERR_INVALID_ARG;
    }
    for (int f = 0; f < n_faces; ++f) {
        if (faces[f].n_vertices < 3 || faces[f].n_vertices > 4 || faces[f].n_inner < 0 || faces[f].n_inner > 4) {
            tif_set_error("tif_plan_create: face %d has %d vertices / %d inner vertices", f, faces[f].n_vertices, faces[f].n_inner);
            return TIF_ERR_INVALID_ARG;
        }
    }
    // the execute kernels address the images of a batch chunk (TIF_MAX_CHUNK_IMAGES images of H*W*3 bytes) with 32-bit
    // byte offsets, and the sample list packs an ERP row / column into 16 bits each
    if ((int64_t)erp_h * erp_w * 3 * TIF_MAX_CHUNK_IMAGES >= (int64_t)1 << 32 || erp_h > 65535 || erp_w > 65535) {
        tif_set_error("tif_plan_create: ERP of %d x %d is beyond the supported size (%d * H * W * 3 bytes must stay below 2^32)",
                      erp_w, erp_h, TIF_MAX_CHUNK_IMAGES);
        return TIF_ERR_UNSUPPORTED;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    int64_t n_tangent = 0;
    for (int f = 0; f < n_faces; ++f) {
        if (faces[f].tangent_h < 2 || faces[f].tangent_h > TIF_MAX_TANGENT_DIM ||
            faces[f].pix_offset != n_tangent) {
            tif_set_error("tif_plan_create: face %d has tangent_h %d / pix_offset %d (expected offset %lld)", f,
                          faces[f].tangent_h, faces[f].pix_offset, (long long)n_tangent);
            return TIF_ERR_INVALID_ARG;
        }
        n_tangent += (int64_t)faces[f].tangent_h * tangent_w;
    }
    if (n_tangent >= (int64_t)1 << 29) {           // 32-bit byte offsets into a chunk of float2 flows
        tif_set_error("tif_plan_create: tangent stack larger than 2^29 pixels");
        return TIF_ERR_UNSUPPORTED;
    }
    const int64_t n_rows = n_tangent / tangent_w;
    const int64_t n_erp = (int64_t)erp_h * erp_w;

    TifPlan* plan = (TifPlan*)calloc(1, sizeof(TifPlan));
    if (!plan) return TIF_ERR_ALLOC;
    plan->n_faces = n_faces;
    plan->erp_h = erp_h;
    plan->erp_w = erp_w;
    plan->tangent_w = tangent_w;
    plan->tangent_pixels = n_tangent;
    plan->kind = kind;
    plan->h_faces = (TifFace*)malloc(sizeof(TifFace) * n_faces);
    if (!plan->h_faces) {
        free(plan);
        return TIF_ERR_ALLOC;
    }
    memcpy(plan->h_faces, faces, sizeof(TifFace) * n_faces);

    PlanScratch sc;          // freed by its destructor on every return below
    std::vector<int32_t> h_row_face((size_t)n_rows), h_first_row((size_t)n_faces);
    {
        int64_t row = 0;
        for (int f = 0; f < n_faces; ++f) {
            h_first_row[f] = (int32_t)row;
            for (int i = 0; i < faces[f].tangent_h; ++i) h_row_face[row++] = f;
        }
    }

    PLAN_TRY(dev_alloc(&plan->d_faces, n_faces, plan));
    PLAN_TRY(dev_alloc(&plan->d_lift_faces, n_faces, plan));
    PLAN_TRY(dev_alloc(&plan->d_pix_face, n_rows, plan));
    PLAN_TRY(dev_alloc(&plan->d_s1_base, n_tangent, plan));
    PLAN_TRY(dev_alloc(&plan->d_s1_frac64, n_tangent, plan));
    PLAN_TRY(dev_alloc(&plan->d_s1_ex, n_tangent, plan));
    PLAN_TRY(dev_alloc(&plan->d_s1_ey, n_tangent, plan));
    PLAN_TRY(dev_alloc(&plan->d_s2_count, n_erp, plan));
    PLAN_TRY(dev_alloc(&plan->d_face_mult_sum, n_faces, plan));
    PLAN_TRY(dev_alloc(&plan->d_ref_mark, n_tangent, plan));
    uint8_t* d_mark = plan->d_ref_mark;
    PLAN_TRY(scratch_alloc(&sc.gx, (size_t)n_faces * tangent_w));
    PLAN_TRY(scratch_alloc(&sc.gy, (size_t)n_rows));
    PLAN_TRY(scratch_alloc(&sc.rows, (size_t)erp_h * 2));
    PLAN_TRY(scratch_alloc(&sc.cols, (size_t)n_faces * erp_w * 2));
    PLAN_TRY(scratch_alloc(&sc.status, 2));
    PLAN_TRY(scratch_alloc(&sc.msum, (size_t)n_faces + 1));

    PLAN_TRY(cuda_status(cudaMemcpyAsync(plan->d_faces, faces, sizeof(TifFace) * n_faces, cudaMemcpyHostToDevice, stream), "copy faces"));
    {
        std::vector<LiftFace> h_lift((size_t)n_faces);
        for (int f = 0; f < n_faces; ++f) {
            LiftFace lf;
            lf.sin_phi0 = faces[f].sin_phi0;
            lf.cos_phi0 = faces[f].cos_phi0;
            lf.xmin = faces[f].xmin;
            lf.ymax = faces[f].ymax;
            lf.kx = 1.0 / faces[f].p2g_x;
            lf.ky = 1.0 / faces[f].p2g_y;
            lf.first_row = faces[f].pix_offset / tangent_w;
            lf.pad[0] = lf.pad[1] = lf.pad[2] = 0;
            h_lift[f] = lf;
        }
        PLAN_TRY(cuda_status(cudaMemcpy(plan->d_lift_faces, h_lift.data(), sizeof(LiftFace) * n_faces, cudaMemcpyHostToDevice), "copy lift faces"));
    }
    PLAN_TRY(cuda_status(cudaMemcpyAsync(plan->d_pix_face, h_row_face.data(), sizeof(int32_t) * n_rows, cudaMemcpyHostToDevice, stream), "copy row->face"));
    PLAN_TRY(cuda_status(cudaMemcpyAsync(sc.gx, tables->gnom_x, sizeof(double) * n_faces * tangent_w, cudaMemcpyHostToDevice, stream), "copy gnom_x"));
    PLAN_TRY(cuda_status(cudaMemcpyAsync(sc.gy, tables->gnom_y, sizeof(double) * n_rows, cudaMemcpyHostToDevice, stream), "copy gnom_y"));
    PLAN_TRY(cuda_status(cudaMemcpyAsync(sc.rows, tables->row_sincos, sizeof(double) * erp_h * 2, cudaMemcpyHostToDevice, stream), "copy row table"));
    PLAN_TRY(cuda_status(cudaMemcpyAsync(sc.cols, tables->col_sincos, sizeof(double) * n_faces * erp_w * 2, cudaMemcpyHostToDevice, stream), "copy column table"));
    PLAN_TRY(cuda_status(cudaMemsetAsync(d_mark, 0, n_tangent, stream), "memset"));
    PLAN_TRY(cuda_status(cudaMemsetAsync(sc.status, 0, 2 * sizeof(int), stream), "memset"));
    PLAN_TRY(cuda_status(cudaMemsetAsync(sc.msum, 0, (n_faces + 1) * sizeof(unsigned long long), stream), "memset"));
    PLAN_TRY(cuda_status(cudaStreamSynchronize(stream), "plan upload"));   // the host tables are pageable

    const int threads = 256;
    k_plan_stage1<<<(unsigned)((n_tangent + threads - 1) / threads), threads, 0, stream>>>(
        plan->d_faces, plan->d_pix_face, sc.gx, sc.gy, erp_h, erp_w, tangent_w, n_tangent, kind, plan->d_s1_base,
        plan->d_s1_frac64, plan->d_s1_ex, plan->d_s1_ey);
    PLAN_TRY(cuda_status(cudaGetLastError(), "k_plan_stage1 launch"));

    const unsigned blocks2 = (unsigned)((n_erp + threads - 1) / threads);
    k_plan_stage2<false><<<blocks2, threads, 0, stream>>>(plan->d_faces, n_faces, erp_h, erp_w, tangent_w, kind, sc.rows,
                                                         sc.cols, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, sc.status);
    PLAN_TRY(cuda_status(cudaGetLastError(), "k_plan_stage2<count> launch"));
    int h_status[2] = {0, 0};
    PLAN_TRY(cuda_status(cudaMemcpyAsync(h_status, sc.status, sizeof(h_status), cudaMemcpyDeviceToHost, stream), "read max count"));
    PLAN_TRY(cuda_status(cudaStreamSynchronize(stream), "plan stage 2 count"));
    plan->max_faces_per_pixel = h_status[0] > 0 ? h_status[0] : 1;
    if (plan->max_faces_per_pixel > 255) {
        tif_set_error("tif_plan_create: %d faces cover one pixel (max 255)", plan->max_faces_per_pixel);
        PLAN_TRY(TIF_ERR_RANGE);
    }
    PLAN_TRY(dev_alloc(&plan->d_s2_entries, (size_t)plan->max_faces_per_pixel * n_erp, plan));
    PLAN_TRY(dev_alloc(&plan->d_s2_snap, (size_t)plan->max_faces_per_pixel * n_erp, plan));
    k_plan_stage2<true><<<blocks2, threads, 0, stream>>>(plan->d_faces, n_faces, erp_h, erp_w, tangent_w, kind, sc.rows,
                                                        sc.cols, plan->d_s2_count, plan->d_s2_entries,
                                                        plan->max_faces_per_pixel, d_mark, sc.msum, plan->d_s1_ex, plan->d_s1_ey,
                                                        plan->d_s2_snap, sc.status);
    PLAN_TRY(cuda_status(cudaGetLastError(), "k_plan_stage2<fill> launch"));
    k_count_marks<<<1024, 256, 0, stream>>>(d_mark, n_tangent, sc.msum + n_faces);
    PLAN_TRY(cuda_status(cudaGetLastError(), "k_count_marks launch"));
    {
        std::vector<unsigned long long> h_msum((size_t)n_faces + 1);
        PLAN_TRY(cuda_status(cudaMemcpyAsync(h_msum.data(), sc.msum, sizeof(unsigned long long) * (n_faces + 1), cudaMemcpyDeviceToHost, stream), "read face sums"));
        PLAN_TRY(cuda_status(cudaMemcpyAsync(h_status, sc.status, sizeof(h_status), cudaMemcpyDeviceToHost, stream), "read status"));
        PLAN_TRY(cuda_status(cudaStreamSynchronize(stream), "plan stage 2 fill"));
        std::vector<double> h_fm((size_t)n_faces);
        for (int f = 0; f < n_faces; ++f) h_fm[f] = (double)h_msum[f];
        plan->referenced_tangent_pixels = (int64_t)h_msum[n_faces];
        PLAN_TRY(cuda_status(cudaMemcpy(plan->d_face_mult_sum, h_fm.data(), sizeof(double) * n_faces, cudaMemcpyHostToDevice), "copy face multiplicity sums"));
    }
    if (h_status[1] != 0) {
        tif_set_error("tif_plan_create: %d membership entries out of range (tangent index outside the face or multiplicity > 7)", h_status[1]);
        PLAN_TRY(TIF_ERR_RANGE);
    }
    // accepted_unique = sum of the per-pixel counts
    PLAN_TRY(cuda_status(cudaMemsetAsync(sc.msum, 0, sizeof(unsigned long long), stream), "memset"));
    k_count_marks<<<1024, 256, 0, stream>>>(plan->d_s2_count, n_erp, sc.msum);
    unsigned long long h_acc = 0;
    PLAN_TRY(cuda_status(cudaMemcpyAsync(&h_acc, sc.msum, sizeof(h_acc), cudaMemcpyDeviceToHost, stream), "read accepted"));
    PLAN_TRY(cuda_status(cudaStreamSynchronize(stream), "plan finish"));
    plan->accepted_unique = (int64_t)h_acc;
    const int64_t n_ref = plan->referenced_tangent_pixels;
    std::vector<int32_t> h_ref_index;           // tangent pixel -> position in the work list, -1 = unreferenced
    {
        // work list of the lift kernels: referenced tangent pixels in 8 x 4 tile order, so that a warp works on a
        // compact patch of one face (plan-time bookkeeping through the host: 1 byte per tangent pixel).
        // An entry names its pixel as face << 25 | row inside the face << 11 | column (the layout of the membership
        // entries), so that the kernel needs neither a division by the tangent width nor the row -> face table.
        std::vector<uint8_t> h_mark((size_t)n_tangent);
        std::vector<int32_t> h_list((size_t)n_ref + 1);
        h_ref_index.assign((size_t)n_tangent, -1);
        PLAN_TRY(cuda_status(cudaMemcpy(h_mark.data(), plan->d_ref_mark, (size_t)n_tangent, cudaMemcpyDeviceToHost), "read referenced marks"));
        int64_t n_list = 0;
        for (int64_t ty = 0; ty < n_rows; ty += 4)
            for (int tx = 0; tx < tangent_w; tx += 8)
                for (int64_t y = ty; y < ty + 4 && y < n_rows; ++y)
                    for (int x = tx; x < tx + 8 && x < tangent_w; ++x)
                        if (h_mark[y * tangent_w + x] && n_list < n_ref) {
                            const int f = h_row_face[y];
                            h_ref_index[y * tangent_w + x] = (int32_t)n_list;
                            h_list[n_list++] = (int32_t)(((uint32_t)f << 25) | ((uint32_t)(y - h_first_row[f]) << 11) | (uint32_t)x);
                        }
        PLAN_TRY(dev_alloc(&plan->d_ref_list, (size_t)n_ref + 1, plan));
        PLAN_TRY(cuda_status(cudaMemcpy(plan->d_ref_list, h_list.data(), sizeof(int32_t) * (size_t)n_list, cudaMemcpyHostToDevice), "referenced pixel list"));
        if (n_list != n_ref) {
            tif_set_error("tif_plan_create: referenced pixel list (%lld) disagrees with the mark count (%lld)",
                          (long long)n_list, (long long)n_ref);
            PLAN_TRY(TIF_ERR_RANGE);
        }
    }
    if (kind == TIF_PLAN_ICOSAHEDRON && n_ref > 0) {
        // normwarp: own ERP position of every work-list entry, and the accepted samples grouped by work-list entry
        PLAN_TRY(dev_alloc(&plan->d_ref_pos, (size_t)n_ref, plan));
        k_plan_ref_pos<<<(unsigned)((n_ref + 255) / 256), 256, 0, stream>>>(plan->d_ref_list, (int)n_ref, plan->d_lift_faces, tangent_w,
                                                                         plan->d_s1_ex, plan->d_s1_ey, plan->d_ref_pos);
        PLAN_TRY(cuda_status(cudaGetLastError(), "k_plan_ref_pos launch"));
        const int64_t n_smp = plan->accepted_unique;
        const int64_t n_blk = (n_ref + TIF_NW_BLOCK - 1) / TIF_NW_BLOCK;
        if (n_smp >= (int64_t)1 << 32) {
            tif_set_error("tif_plan_create: %lld accepted samples (max 2^32)", (long long)n_smp);
            PLAN_TRY(TIF_ERR_UNSUPPORTED);
        }
        PLAN_TRY(scratch_alloc(&sc.ref_index, (size_t)n_tangent));
        PLAN_TRY(scratch_alloc(&sc.per_ref, (size_t)n_ref));
        PLAN_TRY(scratch_alloc(&sc.smp_start, (size_t)n_ref));
        PLAN_TRY(dev_alloc(&plan->d_smp, (size_t)n_smp + 1, plan));
        PLAN_TRY(dev_alloc(&plan->d_blk_smp, (size_t)n_blk + 1, plan));
        PLAN_TRY(cuda_status(cudaMemcpy(sc.ref_index, h_ref_index.data(), sizeof(int32_t) * (size_t)n_tangent, cudaMemcpyHostToDevice), "copy work-list index"));
        PLAN_TRY(cuda_status(cudaMemsetAsync(sc.per_ref, 0, sizeof(uint32_t) * (size_t)n_ref, stream), "memset"));
        k_plan_samples<false><<<blocks2, threads, 0, stream>>>(plan->d_s2_count, plan->d_s2_entries, plan->d_faces, sc.ref_index, erp_h,
                                                              erp_w, tangent_w, sc.per_ref, nullptr, nullptr);
        PLAN_TRY(cuda_status(cudaGetLastError(), "k_plan_samples<count> launch"));
        std::vector<uint32_t> h_cnt((size_t)n_ref), h_blk((size_t)n_blk + 1);
        PLAN_TRY(cuda_status(cudaMemcpyAsync(h_cnt.data(), sc.per_ref, sizeof(uint32_t) * (size_t)n_ref, cudaMemcpyDeviceToHost, stream), "read sample counts"));
        PLAN_TRY(cuda_status(cudaStreamSynchronize(stream), "sample counts"));
        uint64_t run = 0;
        for (int64_t i = 0; i < n_ref; ++i) {        // exclusive scan in work-list order
            if (i % TIF_NW_BLOCK == 0) h_blk[(size_t)(i / TIF_NW_BLOCK)] = (uint32_t)run;
            const uint32_t c = h_cnt[(size_t)i];
            h_cnt[(size_t)i] = (uint32_t)run;
            run += c;
        }
        h_blk[(size_t)n_blk] = (uint32_t)run;
        if ((int64_t)run != n_smp) {
            tif_set_error("tif_plan_create: sample list (%llu) disagrees with the accepted count (%lld)", (unsigned long long)run, (long long)n_smp);
            PLAN_TRY(TIF_ERR_RANGE);
        }
        PLAN_TRY(cuda_status(cudaMemcpy(sc.smp_start, h_cnt.data(), sizeof(uint32_t) * (size_t)n_ref, cudaMemcpyHostToDevice), "copy sample offsets"));
        PLAN_TRY(cuda_status(cudaMemcpy(plan->d_blk_smp, h_blk.data(), sizeof(uint32_t) * ((size_t)n_blk + 1), cudaMemcpyHostToDevice), "copy block sample ranges"));
        PLAN_TRY(cuda_status(cudaMemsetAsync(sc.per_ref, 0, sizeof(uint32_t) * (size_t)n_ref, stream), "memset"));
        k_plan_samples<true><<<blocks2, threads, 0, stream>>>(plan->d_s2_count, plan->d_s2_entries, plan->d_faces, sc.ref_index, erp_h,
                                                             erp_w, tangent_w, sc.per_ref, sc.smp_start, plan->d_smp);
        PLAN_TRY(cuda_status(cudaGetLastError(), "k_plan_samples<fill> launch"));
        PLAN_TRY(cuda_status(cudaStreamSynchronize(stream), "sample list"));
        plan->n_samples = n_smp;
    }
    {
        // stage-1 tile table: the bounding box of the taps of every 16 x 8 tangent patch.  Patches whose box fits a
        // shared-memory stage are resampled from a staged copy of the box; the rest (polar caps, the seam, patches
        // spanning two faces) keep the direct gathers.
        const int tile_w = TIF_S1_TILE_W, tile_h = TIF_S1_TILE_H;
        const int tiles_x = (tangent_w + tile_w - 1) / tile_w, tiles_y = (int)((n_rows + tile_h - 1) / tile_h);
        const size_t n_tiles = (size_t)tiles_x * tiles_y;
        const uint32_t row_bytes = (uint32_t)erp_w * 3u;
        const bool alignable = row_bytes % 16u == 0 && ((size_t)erp_h * row_bytes) % 16u == 0;
        std::vector<uint32_t> h_base((size_t)n_tangent), h_need(n_tiles);     // h_need: bytes of a stage the box needs, 0 = unusable
        std::vector<uint2> h_tiles(n_tiles);
        PLAN_TRY(cuda_status(cudaMemcpy(h_base.data(), plan->d_s1_base, sizeof(uint32_t) * (size_t)n_tangent, cudaMemcpyDeviceToHost), "read stage-1 taps"));
        int64_t staged = 0;
        int stage_bytes = TIF_S1_MIN_STAGE_BYTES;
        for (int ty = 0; ty < tiles_y; ++ty)
            for (int tx = 0; tx < tiles_x; ++tx) {
                uint32_t x_lo = 0xffffffffu, x_hi = 0, y_lo = 0xffffffffu, y_hi = 0;
                for (int64_t y = (int64_t)ty * tile_h; y < (int64_t)(ty + 1) * tile_h && y < n_rows; ++y)
                    for (int x = tx * tile_w; x < (tx + 1) * tile_w && x < tangent_w; ++x) {
                        const uint32_t b = h_base[y * tangent_w + x] & 0x7fffffffu;
                        const uint32_t y0 = b / (uint32_t)erp_w, x0 = b - y0 * (uint32_t)erp_w;
                        x_lo = x0 < x_lo ? x0 : x_lo;
                        x_hi = x0 > x_hi ? x0 : x_hi;
                        y_lo = y0 < y_lo ? y0 : y_lo;
                        y_hi = y0 > y_hi ? y0 : y_hi;
                    }
                uint2 d = make_uint2(0u, 0u);
                uint32_t need = 0u;
                if (alignable && x_lo <= x_hi) {
                    const uint32_t xb0 = (x_lo * 3u) & ~15u;
                    uint32_t xb1 = ((x_hi + 2u) * 3u + 15u) & ~15u;       // taps x0, x0 + 1: 6 bytes from x0 * 3
                    if (xb1 > row_bytes) xb1 = row_bytes;
                    const uint32_t pitch = xb1 - xb0, rows = y_hi + 2u - y_lo;
                    // 16 bytes of slack: the aligned 12-byte window of the last tap may end past its row
                    const size_t bytes = (size_t)rows * pitch + 16u;
                    if (bytes <= TIF_S1_MAX_STAGE_BYTES && y_hi + 1u < (uint32_t)erp_h) {
                        d = make_uint2(y_lo * row_bytes + xb0, rows | (pitch << 16));
                        need = (uint32_t)bytes;
                    }
                }
                h_tiles[(size_t)ty * tiles_x + tx] = d;
                h_need[(size_t)ty * tiles_x + tx] = need;
            }
        // the smallest stage that holds most boxes (small stages leave more shared memory to the L1 and more
        // stages in the ring); what does not fit is resampled by direct gathers
        for (;;) {
            staged = 0;
            for (size_t i = 0; i < n_tiles; ++i) staged += h_need[i] != 0u && h_need[i] <= (uint32_t)stage_bytes;
            if ((double)staged >= TIF_S1_STAGED_SHARE * (double)n_tiles || stage_bytes >= TIF_S1_MAX_STAGE_BYTES) break;
            stage_bytes *= 2;
        }
        for (size_t i = 0; i < n_tiles; ++i)
            if (h_need[i] == 0u || h_need[i] > (uint32_t)stage_bytes) h_tiles[i] = make_uint2(0u, 0u);
        plan->s1_stage_bytes = stage_bytes;
        {
            // Box shapes of the TMA form: a tensor map fixes its box, so two shapes share one row count R (what 95 % of
            // the staged patches need) and differ in the row length: P0 holds about two thirds of them, P1 is the
            // widest that keeps P1 * R inside a stage of TIF_S1_TMA_MAX_STAGE_BYTES (and TMA's 256 elements per box
            // dimension).  A patch outside both keeps the direct gathers in k_erp2ico_tma.
            std::vector<uint32_t> rows_of, pitch_of;
            for (size_t i = 0; i < n_tiles; ++i)
                if (h_tiles[i].y != 0u) {
                    rows_of.push_back(h_tiles[i].y & 0xffffu);
                    pitch_of.push_back(h_tiles[i].y >> 16);
                }
            plan->s1_tma_tiles = 0;
            if (!rows_of.empty()) {
                std::vector<uint32_t> r_sorted(rows_of), p_sorted(pitch_of);
                std::sort(r_sorted.begin(), r_sorted.end());
                std::sort(p_sorted.begin(), p_sorted.end());
                uint32_t box_rows = (r_sorted[(size_t)(0.95 * (r_sorted.size() - 1))] + 3u) & ~3u;
                if (box_rows > 64u) box_rows = 64u;
                uint32_t p1 = ((TIF_S1_TMA_MAX_STAGE_BYTES - 16u) / box_rows) & ~15u;
                if (p1 > 256u) p1 = 256u;
                uint32_t p0 = (p_sorted[(size_t)(0.65 * (p_sorted.size() - 1))] + 15u) & ~15u;
                if (p0 > p1) p0 = p1;
                if (p1 >= 48u) {
                    plan->s1_tma_rows = (int)box_rows;
                    plan->s1_tma_pitch[0] = (int)p0;
                    plan->s1_tma_pitch[1] = (int)p1;
                    for (size_t i = 0; i < rows_of.size(); ++i) plan->s1_tma_tiles += rows_of[i] <= box_rows && pitch_of[i] <= p1;
                }
            }
        }
        PLAN_TRY(dev_alloc(&plan->d_s1_tiles, n_tiles, plan));
        PLAN_TRY(cuda_status(cudaMemcpy(plan->d_s1_tiles, h_tiles.data(), sizeof(uint2) * n_tiles, cudaMemcpyHostToDevice), "stage-1 tile table"));
        plan->s1_tiles_x = tiles_x;
        plan->s1_tiles_y = tiles_y;
        plan->s1_staged_tiles = staged;
    }
    *plan_out = plan;
    return TIF_OK;
}

extern "C" int tif_plan_info(const TifPlan* plan, TifPlanInfo* info) {
    if (!plan || !info) return TIF_ERR_INVALID_ARG;
    info->n_faces = plan->n_faces;
    info->erp_h = plan->erp_h;
    info->erp_w = plan->erp_w;
    info->tangent_w = plan->tangent_w;
    info->tangent_pixels = plan->tangent_pixels;
    info->max_faces_per_pixel = plan->max_faces_per_pixel;
    info->accepted_unique = plan->accepted_unique;
    info->referenced_tangent_pixels = plan->referenced_tangent_pixels;
    info->device_bytes = plan->device_bytes;
    info->stage1_patches = (int64_t)plan->s1_tiles_x * plan->s1_tiles_y;
    info->stage1_staged_patches = plan->s1_staged_tiles;
    info->stage1_tma_patches = plan->s1_tma_tiles;
    info->stage1_tma_box_rows = plan->s1_tma_rows;
    info->stage1_tma_box_bytes[0] = plan->s1_tma_pitch[0];
    info->stage1_tma_box_bytes[1] = plan->s1_tma_pitch[1];
    return TIF_OK;
}

__global__ void k_export_inside(const uint32_t* __restrict__ base, int64_t n, uint8_t* __restrict__ inside) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) inside[p] = (base[p] & TIF_S1_OUTSIDE) ? 0 : 1;
}

extern "C" int tif_plan_export_stage1(const TifPlan* plan, double* erp_x, double* erp_y, uint8_t* inside, void* stream_v) {
    if (!plan) return TIF_ERR_INVALID_ARG;
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int64_t n = plan->tangent_pixels;
    if (erp_x) TIF_CUDA_CHECK(cudaMemcpyAsync(erp_x, plan->d_s1_ex, n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    if (erp_y) TIF_CUDA_CHECK(cudaMemcpyAsync(erp_y, plan->d_s1_ey, n * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    if (inside) {
        k_export_inside<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(plan->d_s1_base, n, inside);
        TIF_CUDA_CHECK(cudaGetLastError());
    }
    return TIF_OK;
}

extern "C" int tif_plan_export_stage2(const TifPlan* plan, uint8_t* count, uint32_t* entries, void* stream_v) {
    if (!plan) return TIF_ERR_INVALID_ARG;
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int64_t n = (int64_t)plan->erp_h * plan->erp_w;
    if (count) TIF_CUDA_CHECK(cudaMemcpyAsync(count, plan->d_s2_count, n, cudaMemcpyDeviceToDevice, stream));
    if (entries)
        TIF_CUDA_CHECK(cudaMemcpyAsync(entries, plan->d_s2_entries, (size_t)plan->max_faces_per_pixel * n * sizeof(uint32_t),
                                       cudaMemcpyDeviceToDevice, stream));
    return TIF_OK;
}
