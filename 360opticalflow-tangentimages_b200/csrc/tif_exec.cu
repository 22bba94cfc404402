// tif_exec.cu -- per-batch execute kernels of libtif_b200.so (sm_100a).
//
// Everything here is a gather / stencil over HBM-resident data: no dense contraction, hence no
// tensor cores.  Geometry comes from the TifPlan (tif_plan.cu); one thread owns one output pixel
// and loops over a chunk of the batch so that plan entries are read once per chunk.
//
//   k_erp2ico        erp2ico_image      projection_icosahedron.py:233-245 (+ scipy map_coordinates, a18)
//   k_stitch         ico2erp_flow       projection_icosahedron.py:333-396, projection.py:15-61
//   k_warp_backward  warp_backward      flow_warp.py:54-88
//   k_meshgrid       flow_warp_meshgrid flow_warp.py:15-51
//   k_wraparound     erp_of_wraparound  flow_postproc.py:17-44
#include "tif_common.cuh"

#define TIF_BATCH_CHUNK 4          // images per thread (plan entries amortised over the chunk)

// ------------------------------------------------------------------------------------------------
// bilinear with scipy's uint8 rounding
// ------------------------------------------------------------------------------------------------
// scipy accumulates p*wy*wx over the taps (y0,x0),(y0,x1),(y1,x0),(y1,x1) in double and stores
// uint8 as trunc(t + 0.5).  The fast path evaluates t in float32; its error is < 2e-4 for 8-bit
// data, so the rounding is already decided unless t + 0.5 lies within GUARD of an integer.  Only
// then (about 1e-3 of samples) is the exact float64 expression evaluated.
#define TIF_ROUND_GUARD 5.0e-4f

__device__ __forceinline__ float bilinear_f32(float p00, float p01, float p10, float p11, float wx0, float wx1,
                                              float wy0, float wy1) {
    return (p00 * wy0) * wx0 + (p01 * wy0) * wx1 + (p10 * wy1) * wx0 + (p11 * wy1) * wx1;
}

__device__ __noinline__ double bilinear_f64_exact(double p00, double p01, double p10, double p11, double fx, double fy) {
    // scipy get_spline_interpolation_weights(order=1): w0 = 1 - f ; w1 = 1 - w0.  No FMA contraction.
    const double wx0 = __dsub_rn(1.0, fx), wx1 = __dsub_rn(1.0, wx0);
    const double wy0 = __dsub_rn(1.0, fy), wy1 = __dsub_rn(1.0, wy0);
    double t = __dmul_rn(__dmul_rn(p00, wy0), wx0);
    t = __dadd_rn(t, __dmul_rn(__dmul_rn(p01, wy0), wx1));
    t = __dadd_rn(t, __dmul_rn(__dmul_rn(p10, wy1), wx0));
    t = __dadd_rn(t, __dmul_rn(__dmul_rn(p11, wy1), wx1));
    return t;
}

__device__ __forceinline__ uint8_t scipy_round_u8(double t) {
    t = t > 0.0 ? t + 0.5 : 0.0;
    t = t > 255.0 ? 255.0 : t;
    return (uint8_t)t;
}

__device__ __forceinline__ uint8_t round_u8_guarded(float t, float p00, float p01, float p10, float p11, double fx,
                                                    double fy) {
    const float r = t + 0.5f;
    const float fl = floorf(r);
    const float fr = r - fl;
    if (fr < TIF_ROUND_GUARD || fr > 1.0f - TIF_ROUND_GUARD)
        return scipy_round_u8(bilinear_f64_exact((double)p00, (double)p01, (double)p10, (double)p11, fx, fy));
    return (uint8_t)fminf(fmaxf(fl, 0.0f), 255.0f);
}

// ------------------------------------------------------------------------------------------------
// stage 1: ERP -> tangent images
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_erp2ico(const uint32_t* __restrict__ s1_base, const float2* __restrict__ s1_frac,
          const double2* __restrict__ s1_frac64, const uint8_t* __restrict__ erp, uint8_t* __restrict__ out,
          int64_t n_pix, int erp_h, int erp_w, int batch, int full_face) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pix) return;
    const int b0 = blockIdx.y * TIF_BATCH_CHUNK;
    const uint32_t bs = s1_base[p];
    const bool outside = (bs & TIF_S1_OUTSIDE) && !full_face;
    const size_t image_bytes = (size_t)erp_h * erp_w * 3;
    const size_t row_bytes = (size_t)erp_w * 3;
    uchar4* out4 = reinterpret_cast<uchar4*>(out);
    if (outside) {
        for (int bb = 0; bb < TIF_BATCH_CHUNK && b0 + bb < batch; ++bb)
            out4[(size_t)(b0 + bb) * n_pix + p] = make_uchar4(255, 255, 255, 0);
        return;
    }
    const float2 fr = s1_frac[p];
    const float wx1 = fr.x, wx0 = 1.0f - fr.x, wy1 = fr.y, wy0 = 1.0f - fr.y;
    const size_t tap = (size_t)(bs & 0x7fffffffu) * 3;
#pragma unroll
    for (int bb = 0; bb < TIF_BATCH_CHUNK; ++bb) {
        const int b = b0 + bb;
        if (b >= batch) break;
        const uint8_t* r0 = erp + (size_t)b * image_bytes + tap;
        const uint8_t* r1 = r0 + row_bytes;
        uint8_t px[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float p00 = (float)__ldg(r0 + ch), p01 = (float)__ldg(r0 + 3 + ch);
            const float p10 = (float)__ldg(r1 + ch), p11 = (float)__ldg(r1 + 3 + ch);
            const float t = bilinear_f32(p00, p01, p10, p11, wx0, wx1, wy0, wy1);
            const float r = t + 0.5f;
            const float fl = floorf(r);
            const float frac = r - fl;
            if (frac < TIF_ROUND_GUARD || frac > 1.0f - TIF_ROUND_GUARD) {
                const double2 f64 = s1_frac64[p];
                px[ch] = scipy_round_u8(bilinear_f64_exact((double)p00, (double)p01, (double)p10, (double)p11, f64.x, f64.y));
            } else {
                px[ch] = (uint8_t)fl;
            }
        }
        out4[(size_t)b * n_pix + p] = make_uchar4(px[0], px[1], px[2], 255);
    }
}

extern "C" int tif_erp2ico_u8(const TifPlan* plan, const uint8_t* erp, int batch, int full_face_image,
                              uint8_t* tangent_rgba, void* stream) {
    if (!plan || !erp || !tangent_rgba || batch <= 0) {
        tif_set_error("tif_erp2ico_u8: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    const int threads = 256;
    dim3 grid((unsigned)((plan->tangent_pixels + threads - 1) / threads), (unsigned)((batch + TIF_BATCH_CHUNK - 1) / TIF_BATCH_CHUNK));
    k_erp2ico<<<grid, threads, 0, (cudaStream_t)stream>>>(plan->d_s1_base, plan->d_s1_frac, plan->d_s1_frac64, erp,
                                                          tangent_rgba, plan->tangent_pixels, plan->erp_h, plan->erp_w,
                                                          batch, full_face_image);
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

// ------------------------------------------------------------------------------------------------
// stage 2+3: tangent flows -> ERP flow
// ------------------------------------------------------------------------------------------------
struct FaceConst {          // per-face constants of the end-point lift, staged in shared memory
    double theta0, sin_phi0, cos_phi0;
    double xmin, ymax;
    double inv_p2g_x, inv_p2g_y;
    int pix_offset;
    int pad;
};

// atan(q) for q in [0,1]: q + q*s*P(s), s = q*q; max abs error 3e-8 (float32 Horner, double final add)
__device__ __forceinline__ double atan_unit(float q) {
    const float s = q * q;
    float r = -1.602407545e-03f;
    r = fmaf(r, s, 1.002551895e-02f);
    r = fmaf(r, s, -2.944629826e-02f);
    r = fmaf(r, s, 5.612712353e-02f);
    r = fmaf(r, s, -8.289746195e-02f);
    r = fmaf(r, s, 1.091027334e-01f);
    r = fmaf(r, s, -1.425546259e-01f);
    r = fmaf(r, s, 1.999762058e-01f);
    r = fmaf(r, s, -3.333326280e-01f);
    return (double)q + (double)(r * (q * s));
}

// atan2(y, x) in (-pi, pi]; float32 ratio + polynomial, double quadrant assembly
__device__ __forceinline__ double atan2_mixed(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float q = mx > 0.0f ? mn / mx : 0.0f;
    double r = atan_unit(q);
    if (ay > ax) r = TIF_HALF_PI - r;
    if (x < 0.0f) r = TIF_PI - r;
    return y < 0.0f ? -r : r;
}

struct Lifted {
    float fu, fv;     // ERP flow of this face at this pixel
    float tx, ty;     // ERP position of the flow end point (after the seam fix), for the normwarp sample
};

// End point of one accepted sample: projection_icosahedron.py:336-366.
//   pixel2gnomonic (gnomonic_projection.py:186-194), reverse_gnomonic_projection (:76-85) written as
//   theta = theta0 + atan2(gx, cos(phi0) - gy sin(phi0)), colatitude = atan2(hypot(.,.), sin(phi0) + gy cos(phi0))
//   (algebraically identical: sin c = rho cos c; better conditioned at the poles than arcsin),
//   sph_coord_modulo + sph2erp(modulo) (spherical_coordinates.py:237-238,59-60,122-124), seam fix (:350-361).
// The direction vector is formed in float64 (cancellation at the poles), the two arctangents in
// float32 with a float64 quadrant sum: <= 1.5e-7 rad, i.e. < 1e-4 px up to 4096-wide ERP.
__device__ __forceinline__ Lifted lift_endpoint(const FaceConst& F, int ix, int iy, float u, float v, double theta_src,
                                                int r, int c, int erp_h, int erp_w, bool wrap_around) {
    const double px = (double)ix + (double)u;
    const double py = (double)iy + (double)v;
    const double gx = px * F.inv_p2g_x + F.xmin;
    const double gy = F.ymax - py * F.inv_p2g_y;
    const double bz = F.cos_phi0 - gy * F.sin_phi0;
    const double cy = F.sin_phi0 + gy * F.cos_phi0;
    const float a32 = (float)gx, b32 = (float)bz, c32 = (float)cy;
    const double dtheta = atan2_mixed(a32, b32);
    const float h32 = sqrtf(a32 * a32 + b32 * b32);
    const double colat = atan2_mixed(h32, c32);           // [0, pi]
    // theta_t wrapped into [-pi, pi) (np.remainder(theta + pi, 2 pi) - pi)
    double tt = F.theta0 + dtheta;
    tt -= TIF_TWO_PI * floor((tt + TIF_PI) * (1.0 / TIF_TWO_PI));
    double d = tt - theta_src;
    if (wrap_around) {
        if (d > TIF_PI && theta_src < 0.0 && tt > 0.0) d -= TIF_TWO_PI;
        else if (d < -TIF_PI && theta_src > 0.0 && tt < 0.0) d += TIF_TWO_PI;
    }
    Lifted o;
    const double fu = d * ((double)erp_w * (1.0 / TIF_TWO_PI));
    const double ty = colat * ((double)erp_h * (1.0 / TIF_PI)) - 0.5;
    o.fu = (float)fu;
    o.fv = (float)(ty - (double)r);
    o.tx = (float)(fu + (double)c);
    o.ty = (float)ty;
    return o;
}

// mean over channels of |bilinear(tar, (ty,tx)) - src[r,c]| with scipy mode='constant', cval=255
// (projection.py:52-55): any coordinate outside [0, n-1] reads 255 outright.
__device__ __forceinline__ float warp_error_sum(const uint8_t* __restrict__ tar, const uint8_t* __restrict__ src_px,
                                                float tx, float ty, int erp_h, int erp_w) {
    float t0, t1, t2;
    if (tx < 0.0f || tx > (float)(erp_w - 1) || ty < 0.0f || ty > (float)(erp_h - 1)) {
        t0 = t1 = t2 = 255.0f;
    } else {
        float x0 = floorf(tx), y0 = floorf(ty);
        x0 = fminf(x0, (float)(erp_w - 2));
        y0 = fminf(y0, (float)(erp_h - 2));
        const float fx = tx - x0, fy = ty - y0;
        const float wx0 = 1.0f - fx, wy0 = 1.0f - fy;
        const uint8_t* r0 = tar + ((size_t)(int)y0 * erp_w + (int)x0) * 3;
        const uint8_t* r1 = r0 + (size_t)erp_w * 3;
        t0 = bilinear_f32(__ldg(r0 + 0), __ldg(r0 + 3), __ldg(r1 + 0), __ldg(r1 + 3), wx0, fx, wy0, fy);
        t1 = bilinear_f32(__ldg(r0 + 1), __ldg(r0 + 4), __ldg(r1 + 1), __ldg(r1 + 4), wx0, fx, wy0, fy);
        t2 = bilinear_f32(__ldg(r0 + 2), __ldg(r0 + 5), __ldg(r1 + 2), __ldg(r1 + 5), wx0, fx, wy0, fy);
    }
    return fabsf(t0 - (float)src_px[0]) + fabsf(t1 - (float)src_px[1]) + fabsf(t2 - (float)src_px[2]);
}

__device__ __forceinline__ void load_face_consts(FaceConst* sh, const TifFace* __restrict__ faces, int n_faces) {
    for (int f = threadIdx.x; f < n_faces; f += blockDim.x) {
        const TifFace& F = faces[f];
        FaceConst fc;
        fc.theta0 = F.theta0;
        fc.sin_phi0 = F.sin_phi0;
        fc.cos_phi0 = F.cos_phi0;
        fc.xmin = F.xmin;
        fc.ymax = F.ymax;
        fc.inv_p2g_x = 1.0 / F.p2g_x;
        fc.inv_p2g_y = 1.0 / F.p2g_y;
        fc.pix_offset = F.pix_offset;
        fc.pad = 0;
        sh[f] = fc;
    }
}

// MODE 0: straightforward blend.  MODE 1: normwarp pass A (face-global warp-error sums only).
// MODE 2: normwarp pass B (weights from the sums of pass A, blend).
template <int MODE>
__global__ void __launch_bounds__(256)
k_stitch(const TifFace* __restrict__ faces, int n_faces, const uint8_t* __restrict__ s2_count,
         const uint32_t* __restrict__ s2_entries, const double* __restrict__ col_theta,
         const double* __restrict__ face_mult_sum, const float* __restrict__ tflow,
         const uint8_t* __restrict__ erp_src, const uint8_t* __restrict__ erp_tar,
         unsigned long long* __restrict__ nw_sums /* [B][F] fixed point */, float* __restrict__ out,
         int erp_h, int erp_w, int tangent_w, int64_t n_tangent, int batch, int wrap_around) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FaceConst* sh_face = reinterpret_cast<FaceConst*>(smem_raw);
    // MODE 1: per-block partial sums; MODE 2: 1 / face-global mean
    unsigned long long* sh_sum = reinterpret_cast<unsigned long long*>(sh_face + n_faces);
    float* sh_inv_mean = reinterpret_cast<float*>(sh_face + n_faces);

    const int b0 = blockIdx.y * TIF_BATCH_CHUNK;
    load_face_consts(sh_face, faces, n_faces);
    if (MODE == 1) {
        for (int i = threadIdx.x; i < TIF_BATCH_CHUNK * n_faces; i += blockDim.x) sh_sum[i] = 0ull;
    } else if (MODE == 2) {
        for (int i = threadIdx.x; i < TIF_BATCH_CHUNK * n_faces; i += blockDim.x) {
            const int bb = i / n_faces, f = i - bb * n_faces;
            float inv = 0.0f;
            if (b0 + bb < batch) {
                // np.mean over all accepted samples x 3 channels, duplicates counted (projection.py:57)
                const double total = (double)nw_sums[(size_t)(b0 + bb) * n_faces + f] / (double)TIF_NW_FIXED_SCALE;
                const double mean = total / (3.0 * face_mult_sum[f]);
                inv = (float)(1.0 / mean);
            }
            sh_inv_mean[i] = inv;
        }
    }
    __syncthreads();

    const int64_t n_pix = (int64_t)erp_h * erp_w;
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = pix < n_pix;
    int cnt = 0, r = 0, c = 0;
    double theta_src = 0.0;
    if (live) {
        cnt = s2_count[pix];
        r = (int)(pix / erp_w);
        c = (int)(pix - (int64_t)r * erp_w);
        // sph_coord_modulo of the source longitude (projection_icosahedron.py:353)
        theta_src = col_theta[c];
        theta_src -= TIF_TWO_PI * floor((theta_src + TIF_PI) * (1.0 / TIF_TWO_PI));
    }
    float acc_u[TIF_BATCH_CHUNK], acc_v[TIF_BATCH_CHUNK], acc_w[TIF_BATCH_CHUNK];
#pragma unroll
    for (int bb = 0; bb < TIF_BATCH_CHUNK; ++bb) acc_u[bb] = acc_v[bb] = acc_w[bb] = 0.0f;
    const size_t image_bytes = (size_t)n_pix * 3;

    const int max_cnt = (MODE == 1) ? __reduce_max_sync(0xffffffffu, cnt) : cnt;
    for (int k = 0; k < max_cnt; ++k) {
        const bool has = k < cnt;
        uint32_t e = 0;
        if (has) e = __ldg(s2_entries + (int64_t)k * n_pix + pix);
        const int f = (int)TIF_E_FACE(e);
        const FaceConst& F = sh_face[f];
        const int ix = (int)TIF_E_IX(e), iy = (int)TIF_E_IY(e);
        const int64_t tpix = (int64_t)F.pix_offset + (int64_t)iy * tangent_w + ix;
#pragma unroll
        for (int bb = 0; bb < TIF_BATCH_CHUNK; ++bb) {
            const int b = b0 + bb;
            if (b >= batch) break;
            float dsum = 0.0f;
            Lifted L;
            if (has) {
                const float2 uv = __ldg(reinterpret_cast<const float2*>(tflow) + (size_t)b * n_tangent + tpix);
                L = lift_endpoint(F, ix, iy, uv.x, uv.y, theta_src, r, c, erp_h, erp_w, wrap_around != 0);
                if (MODE != 0)
                    dsum = warp_error_sum(erp_tar + (size_t)b * image_bytes, erp_src + (size_t)b * image_bytes + (size_t)pix * 3,
                                          L.tx, L.ty, erp_h, erp_w);
            }
            if (MODE == 0) {
                if (has) {
                    acc_u[bb] += L.fu;
                    acc_v[bb] += L.fv;
                    acc_w[bb] += 1.0f;
                }
            } else if (MODE == 1) {
                // face-global sum of |diff| (x multiplicity) as order-independent fixed point.
                // Neighbouring pixels almost always see the same face at the same k: reduce across
                // the warp first and issue one shared-memory atomic; otherwise one per lane.
                const unsigned fixed = has ? (unsigned)(dsum * (float)TIF_E_MULT(e) * TIF_NW_FIXED_SCALE + 0.5f) : 0u;
                const int f0 = __shfl_sync(0xffffffffu, f, 0);
                const bool uniform = __all_sync(0xffffffffu, (!has) || f == f0) && __shfl_sync(0xffffffffu, (int)has, 0);
                if (uniform) {
                    // 32 lanes x (765 x 7 x 32768) overflows 32 bits: reduce as two 16-bit halves
                    const unsigned lo = __reduce_add_sync(0xffffffffu, fixed & 0xffffu);
                    const unsigned hi = __reduce_add_sync(0xffffffffu, fixed >> 16);
                    if ((threadIdx.x & 31) == 0)
                        atomicAdd(&sh_sum[bb * n_faces + f0], (unsigned long long)lo + ((unsigned long long)hi << 16));
                } else if (has) {
                    atomicAdd(&sh_sum[bb * n_faces + f], (unsigned long long)fixed);
                }
            } else {
                if (has) {
                    // projection.py:57-58: exp(-(mean_ch |diff|) / mean_all)
                    const float w = __expf(-(dsum * (1.0f / 3.0f)) * sh_inv_mean[bb * n_faces + f]);
                    acc_u[bb] += L.fu * w;
                    acc_v[bb] += L.fv * w;
                    acc_w[bb] += w;
                }
            }
        }
    }

    if (MODE == 1) {
        __syncthreads();
        for (int i = threadIdx.x; i < TIF_BATCH_CHUNK * n_faces; i += blockDim.x) {
            const int bb = i / n_faces, f = i - bb * n_faces;
            if (b0 + bb < batch && sh_sum[i]) atomicAdd(&nw_sums[(size_t)(b0 + bb) * n_faces + f], sh_sum[i]);
        }
        return;
    }
    if (!live) return;
    const float half_w = 0.5f * (float)erp_w;
    const int split = (int)((double)erp_w - 1.0 - 0.5 * (double)erp_w);
#pragma unroll
    for (int bb = 0; bb < TIF_BATCH_CHUNK; ++bb) {
        const int b = b0 + bb;
        if (b >= batch) break;
        float u = acc_u[bb], v = acc_v[bb];
        if (acc_w[bb] != 0.0f) {      // projection_icosahedron.py:389-393
            u = u / acc_w[bb];
            v = v / acc_w[bb];
        }
        if (wrap_around) {            // flow_postproc.erp_of_wraparound, flow_postproc.py:28-42
            if (c < split) {
                if (u > half_w) u -= (float)erp_w;
            } else if (c < erp_w - 1) {
                if (u < -half_w) u += (float)erp_w;
            }
        }
        reinterpret_cast<float2*>(out)[(size_t)b * n_pix + pix] = make_float2(u, v);
    }
}

extern "C" int64_t tif_ico2erp_scratch_bytes(const TifPlan* plan, int batch, int blend) {
    if (!plan || batch <= 0) return 0;
    if (blend != TIF_BLEND_NORMWARP) return 0;
    return (int64_t)batch * plan->n_faces * (int64_t)sizeof(unsigned long long);
}

extern "C" int tif_ico2erp_flow_f32_ex(const TifPlan* plan, const float* tangent_flow, const uint8_t* erp_src,
                                       const uint8_t* erp_tar, int batch, int blend, int wrap_around, float* erp_flow,
                                       void* scratch, int pass_mask, void* stream_v) {
    if (!plan || !tangent_flow || !erp_flow || batch <= 0) {
        tif_set_error("tif_ico2erp_flow_f32: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    if (blend != TIF_BLEND_STRAIGHTFORWARD && blend != TIF_BLEND_NORMWARP) {
        tif_set_error("tif_ico2erp_flow_f32: unknown blend %d", blend);
        return TIF_ERR_INVALID_ARG;
    }
    if (blend == TIF_BLEND_NORMWARP && (!erp_src || !erp_tar || !scratch)) {
        tif_set_error("tif_ico2erp_flow_f32: normwarp needs erp_src, erp_tar and scratch");
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int threads = 256;
    const int64_t n_pix = (int64_t)plan->erp_h * plan->erp_w;
    dim3 grid((unsigned)((n_pix + threads - 1) / threads), (unsigned)((batch + TIF_BATCH_CHUNK - 1) / TIF_BATCH_CHUNK));
    const size_t smem = sizeof(FaceConst) * plan->n_faces + sizeof(unsigned long long) * TIF_BATCH_CHUNK * plan->n_faces;
#define STITCH_ARGS                                                                                                 \
    plan->d_faces, plan->n_faces, plan->d_s2_count, plan->d_s2_entries, plan->d_col_theta, plan->d_face_mult_sum,   \
        tangent_flow, erp_src, erp_tar, (unsigned long long*)scratch, erp_flow, plan->erp_h, plan->erp_w,           \
        plan->tangent_w, plan->tangent_pixels, batch, wrap_around
    if (blend == TIF_BLEND_STRAIGHTFORWARD) {
        k_stitch<0><<<grid, threads, smem, stream>>>(STITCH_ARGS);
    } else {
        if (pass_mask & 1) {
            TIF_CUDA_CHECK(cudaMemsetAsync(scratch, 0, (size_t)tif_ico2erp_scratch_bytes(plan, batch, blend), stream));
            k_stitch<1><<<grid, threads, smem, stream>>>(STITCH_ARGS);
            TIF_CUDA_CHECK(cudaGetLastError());
        }
        if (pass_mask & 2) k_stitch<2><<<grid, threads, smem, stream>>>(STITCH_ARGS);
    }
#undef STITCH_ARGS
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

extern "C" int tif_ico2erp_flow_f32(const TifPlan* plan, const float* tangent_flow, const uint8_t* erp_src,
                                    const uint8_t* erp_tar, int batch, int blend, int wrap_around, float* erp_flow,
                                    void* scratch, void* stream) {
    return tif_ico2erp_flow_f32_ex(plan, tangent_flow, erp_src, erp_tar, batch, blend, wrap_around, erp_flow, scratch, 3,
                                   stream);
}

// ------------------------------------------------------------------------------------------------
// flow_warp.warp_backward
// ------------------------------------------------------------------------------------------------
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
    // flow_warp.py:75-76: mesh grid + flow, in float64
    double x = (double)c + (double)flow[2 * gi];
    double y = (double)r + (double)flow[2 * gi + 1];
    ImgT* o = out + gi * channels;
    if (MODE == TIF_WARP_CONSTANT255) {
        if (x < 0.0 || x > (double)(w - 1) || y < 0.0 || y > (double)(h - 1)) {
            for (int ch = 0; ch < channels; ++ch) o[ch] = (ImgT)255;
            return;
        }
    } else {
        x = scipy_wrap(x, w);
        y = scipy_wrap(y, h);
    }
    double x0 = floor(x), y0 = floor(y);
    if (x0 > (double)(w - 2)) x0 = (double)(w - 2);
    if (y0 > (double)(h - 2)) y0 = (double)(h - 2);
    if (x0 < 0.0) x0 = 0.0;
    if (y0 < 0.0) y0 = 0.0;
    const double fx = x - x0, fy = y - y0;
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
        for (int ch = 0; ch < channels; ++ch) {
            const double t = bilinear_f64_exact((double)__ldg(r0 + ch), (double)__ldg(r0 + channels + ch),
                                                (double)__ldg(r1 + ch), (double)__ldg(r1 + channels + ch), fx, fy);
            o[ch] = (ImgT)t;
        }
    }
}

template <typename ImgT>
static int launch_warp(const ImgT* image, const void* flow, int flow_dtype, int batch, int h, int w, int channels,
                       int mode, ImgT* out, void* stream_v) {
    if (!image || !flow || !out || batch <= 0 || h < 2 || w < 2 || channels <= 0) {
        tif_set_error("tif_warp_backward: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int threads = 256;
    dim3 grid((unsigned)(((int64_t)h * w + threads - 1) / threads), (unsigned)batch);
    if (flow_dtype == TIF_DTYPE_F32 && mode == TIF_WARP_WRAP)
        k_warp_backward<ImgT, float, TIF_WARP_WRAP><<<grid, threads, 0, stream>>>(image, (const float*)flow, out, batch, h, w, channels);
    else if (flow_dtype == TIF_DTYPE_F32 && mode == TIF_WARP_CONSTANT255)
        k_warp_backward<ImgT, float, TIF_WARP_CONSTANT255><<<grid, threads, 0, stream>>>(image, (const float*)flow, out, batch, h, w, channels);
    else if (flow_dtype == TIF_DTYPE_F64 && mode == TIF_WARP_WRAP)
        k_warp_backward<ImgT, double, TIF_WARP_WRAP><<<grid, threads, 0, stream>>>(image, (const double*)flow, out, batch, h, w, channels);
    else if (flow_dtype == TIF_DTYPE_F64 && mode == TIF_WARP_CONSTANT255)
        k_warp_backward<ImgT, double, TIF_WARP_CONSTANT255><<<grid, threads, 0, stream>>>(image, (const double*)flow, out, batch, h, w, channels);
    else {
        tif_set_error("tif_warp_backward: unknown flow dtype %d / mode %d", flow_dtype, mode);
        return TIF_ERR_INVALID_ARG;
    }
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
