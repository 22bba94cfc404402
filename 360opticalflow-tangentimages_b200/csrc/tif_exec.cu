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

#ifndef TIF_BATCH_CHUNK
#define TIF_BATCH_CHUNK 4          // images per thread (plan entries amortised over the chunk)
#endif

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
// The two taps of one image row are 6 consecutive bytes (R0 G0 B0 R1 G1 B1) at an arbitrary byte
// address.  They are fetched with two or three aligned 4-byte loads and funnel-shifted into place
// instead of six byte loads (the third word is only needed at byte phase 3).  `o` = address & 3.  The
// caller guarantees that 12 bytes from the aligned address are readable.
struct RawPair {
    uint32_t w0, w1, w2;
};

__device__ __forceinline__ RawPair load_raw_pair(const uint8_t* __restrict__ p, unsigned o) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p - o);
    RawPair r;
    r.w0 = __ldg(q);
    r.w1 = __ldg(q + 1);
    r.w2 = 0u;
    if (o == 3) r.w2 = __ldg(q + 2);
    return r;
}

__device__ __forceinline__ void align_pair(const RawPair& r, unsigned o, uint32_t& lo /* R0 G0 B0 R1 */,
                                           uint32_t& hi /* G1 B1 . . */) {
    const unsigned sh = o * 8;
    lo = __funnelshift_r(r.w0, r.w1, sh);
    hi = __funnelshift_r(r.w1, r.w2, sh);
}

__device__ __forceinline__ void load_tap_pair(const uint8_t* __restrict__ p, unsigned o, uint32_t& lo, uint32_t& hi) {
    const RawPair r = load_raw_pair(p, o);
    align_pair(r, o, lo, hi);
}

__device__ __forceinline__ void load_tap_pair_bytes(const uint8_t* __restrict__ p, uint32_t& lo, uint32_t& hi) {
    lo = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) | ((uint32_t)__ldg(p + 3) << 24);
    hi = (uint32_t)__ldg(p + 4) | ((uint32_t)__ldg(p + 5) << 8);
}

#define TIF_BYTE(w, k) __byte_perm((w), 0u, 0x4440u | (k))       // zero-extended byte k: one PRMT, no XU conversion

// Fixed-point bilinear: weights scaled by 2^24 (exactly rounded from the float64 scipy weights), so the
// top byte of p00 W00 + p01 W01 + p10 W10 + p11 W11 + 2^23 is scipy's trunc(t + 0.5).  The fixed-point
// sum is within 3.1e-5 of t; if t + 0.5 lies within 2^-11 of an integer the exact float64 expression
// decides instead (about 1e-3 of the samples).
#define TIF_FIX_ONE 16777216.0
#define TIF_FIX_HALF 0x00800000u
#define TIF_FIX_GUARD 0x00002000u

__device__ __forceinline__ uint32_t fix_bilinear(uint32_t p00, uint32_t p01, uint32_t p10, uint32_t p11, uint32_t w00,
                                                 uint32_t w01, uint32_t w10, uint32_t w11) {
    return p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + TIF_FIX_HALF;
}

__device__ __forceinline__ bool fix_undecided(uint32_t acc) {
    return ((acc + TIF_FIX_GUARD) & 0x00ffffffu) < 2u * TIF_FIX_GUARD;
}

// one RGBA output pixel from the two aligned tap pairs (R: bytes 0,3 of lo; G: byte 1 of lo, byte 0 of hi;
// B: byte 2 of lo, byte 1 of hi)
__device__ __forceinline__ uint32_t erp2ico_pixel(uint32_t lo0, uint32_t hi0, uint32_t lo1, uint32_t hi1, uint32_t w00,
                                                  uint32_t w01, uint32_t w10, uint32_t w11, const double2& fr) {
    uint32_t ar = fix_bilinear(TIF_BYTE(lo0, 0), TIF_BYTE(lo0, 3), TIF_BYTE(lo1, 0), TIF_BYTE(lo1, 3), w00, w01, w10, w11);
    uint32_t ag = fix_bilinear(TIF_BYTE(lo0, 1), TIF_BYTE(hi0, 0), TIF_BYTE(lo1, 1), TIF_BYTE(hi1, 0), w00, w01, w10, w11);
    uint32_t ab = fix_bilinear(TIF_BYTE(lo0, 2), TIF_BYTE(hi0, 1), TIF_BYTE(lo1, 2), TIF_BYTE(hi1, 1), w00, w01, w10, w11);
    if (fix_undecided(ar) | fix_undecided(ag) | fix_undecided(ab)) {
        ar = (uint32_t)scipy_round_u8(bilinear_f64_exact(TIF_BYTE(lo0, 0), TIF_BYTE(lo0, 3), TIF_BYTE(lo1, 0), TIF_BYTE(lo1, 3), fr.x, fr.y)) << 24;
        ag = (uint32_t)scipy_round_u8(bilinear_f64_exact(TIF_BYTE(lo0, 1), TIF_BYTE(hi0, 0), TIF_BYTE(lo1, 1), TIF_BYTE(hi1, 0), fr.x, fr.y)) << 24;
        ab = (uint32_t)scipy_round_u8(bilinear_f64_exact(TIF_BYTE(lo0, 2), TIF_BYTE(hi0, 1), TIF_BYTE(lo1, 2), TIF_BYTE(hi1, 1), fr.x, fr.y)) << 24;
    }
    // top bytes -> R | G<<8 | B<<16 | 255<<24
    return __byte_perm(__byte_perm(ar, ag, 0x0073u), ab, 0x0710u) | 0xff000000u;
}

// A warp owns an 8 x 4 tile of the tangent stack, not a 32 x 1 strip: tangent rows are tilted against ERP
// rows (steeply on the polar faces), so a strip touches ~21 cache lines per tap load and a tile ~9.
#define TIF_TILE_W 8
#define TIF_TILE_H 4

__device__ __forceinline__ bool tile_pixel(int64_t idx, int width, int rows, int& row, int& col) {
    const int64_t warp = idx >> 5;
    const int lane = (int)(idx & 31);
    const int tiles_x = (width + TIF_TILE_W - 1) / TIF_TILE_W;
    const int ty = (int)(warp / tiles_x), tx = (int)(warp - (int64_t)ty * tiles_x);
    row = ty * TIF_TILE_H + lane / TIF_TILE_W;
    col = tx * TIF_TILE_W + lane % TIF_TILE_W;
    return row < rows && col < width;
}

__host__ __device__ __forceinline__ int64_t tile_threads(int width, int rows) {
    return (int64_t)((width + TIF_TILE_W - 1) / TIF_TILE_W) * ((rows + TIF_TILE_H - 1) / TIF_TILE_H) * 32;
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256)
k_erp2ico(const uint32_t* __restrict__ s1_base, const double2* __restrict__ s1_frac64,
          const uint8_t* __restrict__ erp, uint8_t* __restrict__ out, int64_t n_pix, int tangent_w, int erp_h, int erp_w,
          int batch, int batch_per_block_row, int full_face) {
    int trow, tcol;
    if (!tile_pixel((int64_t)blockIdx.x * blockDim.x + threadIdx.x, tangent_w, (int)(n_pix / tangent_w), trow, tcol)) return;
    const int64_t p = (int64_t)trow * tangent_w + tcol;
    const int b_begin = blockIdx.y * batch_per_block_row;
    const int b_end = min(batch, b_begin + batch_per_block_row);
    const uint32_t bs = s1_base[p];
    const size_t image_bytes = (size_t)erp_h * erp_w * 3;
    const size_t row_bytes = (size_t)erp_w * 3;
    uint32_t* o4 = reinterpret_cast<uint32_t*>(out) + (size_t)b_begin * n_pix + p;
    if ((bs & TIF_S1_OUTSIDE) && !full_face) {
        for (int b = b_begin; b < b_end; ++b, o4 += n_pix) *o4 = 0x00ffffffu;   // RGB 255, alpha 0
        return;
    }
    const double2 fr = s1_frac64[p];
    uint32_t w00, w01, w10, w11;
    {
        // scipy: w0 = 1 - f, w1 = 1 - w0 per axis; tap weight = wy * wx
        const double wx0 = 1.0 - fr.x, wx1 = 1.0 - wx0, wy0 = 1.0 - fr.y, wy1 = 1.0 - wy0;
        w00 = (uint32_t)(wy0 * wx0 * TIF_FIX_ONE + 0.5);
        w01 = (uint32_t)(wy0 * wx1 * TIF_FIX_ONE + 0.5);
        w10 = (uint32_t)(wy1 * wx0 * TIF_FIX_ONE + 0.5);
        w11 = (uint32_t)(wy1 * wx1 * TIF_FIX_ONE + 0.5);
    }
    const size_t tap = (size_t)(bs & 0x7fffffffu) * 3;
    const uint8_t* r0 = erp + (size_t)b_begin * image_bytes + tap;
    // The wide loads may touch up to 15 bytes past the second tap pair; only a tap at the very end of an
    // image could leave the buffer, and only in the last image: those few pixels use byte loads.
    const bool tail = tap + row_bytes + 12 > image_bytes;
    if (tail) {
        for (int b = b_begin; b < b_end; ++b, r0 += image_bytes, o4 += n_pix) {
            uint32_t lo0, hi0, lo1, hi1;
            load_tap_pair_bytes(r0, lo0, hi0);
            load_tap_pair_bytes(r0 + row_bytes, lo1, hi1);
            *o4 = erp2ico_pixel(lo0, hi0, lo1, hi1, w00, w01, w10, w11, fr);
        }
        return;
    }
    unsigned o0 = (unsigned)(reinterpret_cast<uintptr_t>(r0) & 3);
    unsigned o1 = (unsigned)(reinterpret_cast<uintptr_t>(r0 + row_bytes) & 3);
    // the kernel is latency bound (taps are L1/L2 hits, not streams): keep the raw words of the next image in
    // flight while the current one is blended
    RawPair c0 = load_raw_pair(r0, o0), c1 = load_raw_pair(r0 + row_bytes, o1);
    for (int b = b_begin; b < b_end; ++b, o4 += n_pix) {
        RawPair n0 = c0, n1 = c1;
        unsigned no0 = o0, no1 = o1;
        r0 += image_bytes;
        if (b + 1 < b_end) {
            if (!ALIGNED) {     // image stride not a multiple of 4: the byte phase changes per image
                no0 = (unsigned)(reinterpret_cast<uintptr_t>(r0) & 3);
                no1 = (unsigned)(reinterpret_cast<uintptr_t>(r0 + row_bytes) & 3);
            }
            n0 = load_raw_pair(r0, no0);
            n1 = load_raw_pair(r0 + row_bytes, no1);
        }
        uint32_t lo0, hi0, lo1, hi1;
        align_pair(c0, o0, lo0, hi0);
        align_pair(c1, o1, lo1, hi1);
        *o4 = erp2ico_pixel(lo0, hi0, lo1, hi1, w00, w01, w10, w11, fr);
        c0 = n0;
        c1 = n1;
        o0 = no0;
        o1 = no1;
    }
}

extern "C" int tif_erp2ico_u8(const TifPlan* plan, const uint8_t* erp, int batch, int full_face_image,
                              uint8_t* tangent_rgba, void* stream) {
    if (!plan || !erp || !tangent_rgba || batch <= 0) {
        tif_set_error("tif_erp2ico_u8: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    const int threads = 256;
    const int64_t n_threads = tile_threads(plan->tangent_w, (int)(plan->tangent_pixels / plan->tangent_w));
    const unsigned blocks_x = (unsigned)((n_threads + threads - 1) / threads);
    // every thread walks the whole batch so that the plan entry is read once; only tiny problems
    // split the batch over blockIdx.y to fill the 148 SMs
    int rows = 1;
    while (rows < batch && (int64_t)blocks_x * rows < 148 * 16) rows *= 2;
    const int per_row = (batch + rows - 1) / rows;
    dim3 grid(blocks_x, (unsigned)((batch + per_row - 1) / per_row));
    const size_t image_bytes = (size_t)plan->erp_h * plan->erp_w * 3;
    const bool aligned = (image_bytes % 4 == 0) && (reinterpret_cast<uintptr_t>(erp) % 4 == 0);
    if (aligned)
        k_erp2ico<true><<<grid, threads, 0, (cudaStream_t)stream>>>(plan->d_s1_base, plan->d_s1_frac64, erp, tangent_rgba,
                                                                    plan->tangent_pixels, plan->tangent_w, plan->erp_h,
                                                                    plan->erp_w, batch, per_row, full_face_image);
    else
        k_erp2ico<false><<<grid, threads, 0, (cudaStream_t)stream>>>(plan->d_s1_base, plan->d_s1_frac64, erp, tangent_rgba,
                                                                     plan->tangent_pixels, plan->tangent_w, plan->erp_h,
                                                                     plan->erp_w, batch, per_row, full_face_image);
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

// ------------------------------------------------------------------------------------------------
// stage 2+3: tangent flows -> ERP flow
// ------------------------------------------------------------------------------------------------
#ifndef TIF_STITCH_THREADS
#define TIF_STITCH_THREADS 128
#endif
#ifndef TIF_STITCH_MIN_BLOCKS
#define TIF_STITCH_MIN_BLOCKS 8       // 64 registers per thread: occupancy over a few spilled values
#endif
// rows whose sin(colatitude) is below this use the float64 end-point lift: an end point next to the polar
// axis needs the float64 products of the flow (the float32 local frame cancels there), and under normwarp
// the faces disagree by hundreds of ERP pixels next to the poles, so the photometric weights turn an
// end-point error of 1e-4 px into a flow error above the 1e-3 px tolerance
#define TIF_POLAR_SIN 0.1f

struct FaceConst {          // per-face constants of the end-point lift, staged in shared memory
    double sin_phi0, cos_phi0;
    double xmin, ymax;
    double kx, ky;              // 1 / pixel2gnomonic ratios (gnomonic_projection.py:188-194)
    double ky_sin, ky_cos;      // ky * sin(phi0), ky * cos(phi0)
    int pix_offset;
    int pad;
};

// atan(y/x) for x > 0 and |y| <= 0.4143 x: q + q s P(s); max abs error 2.3e-8
// reciprocal for x in the normal range: MUFU.RCP + one Newton step (no denormal / range branches)
__device__ __forceinline__ float rcp_newton(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}

// sqrt for x > 0 in the normal range: MUFU.RSQ + one Newton step
__device__ __forceinline__ float sqrt_newton(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    const float h = x * r;
    return fmaf(0.5f * r, fmaf(-h, h, x), h);
}

__device__ __forceinline__ float atan_small(float y, float x) {
    const float q = y * rcp_newton(x);
    const float s = q * q;
    float r = 7.901539654e-02f;
    r = fmaf(r, s, -1.382412463e-01f);
    r = fmaf(r, s, 1.997184753e-01f);
    r = fmaf(r, s, -3.333275616e-01f);
    return fmaf(r, q * s, q);
}

// atan(q) for q in [0,1]: q + q s P(s); max abs error 3e-8 before the final float32 rounding
__device__ __forceinline__ float atan_unit(float q) {
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
    return fmaf(r, q * s, q);
}

// atan2(y, x) in (-pi, pi].  The common case of the stitch (a small angle between the flow end point
// and its source pixel) takes the short polynomial; everything else the full octant reduction.
__device__ __forceinline__ float atan2_fast(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    if (x > 0.0f && ay <= 0.4143f * x) return atan_small(y, x);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float q = mn * rcp_newton(fmaxf(mx, 1e-30f));
    float r = atan_unit(q);
    if (ay > ax) r = (float)TIF_HALF_PI - r;
    if (x < 0.0f) r = (float)TIF_PI - r;
    return y < 0.0f ? -r : r;
}

struct SampleGeom {       // per (ERP pixel, accepting face): constant over the batch
    double gx0, bz0, cy0;           // direction of the tangent pixel centre: gx, cos(phi0) - gy sin(phi0), sin(phi0) + gy cos(phi0)
    double kx, ky_sin, ky_cos;
    // float32 local frame of the source pixel (bulk rows): the end point is the source direction plus a small
    // displacement (dx, dy) on the tangent plane, decomposed into east / north / radial parts
    float d0x, d0y;                 // tangent-pixel centre minus the source pixel's own gnomonic position
    float kxf, kyf;
    float e1, e2, n1, n2, u1, u2, u0;
    int tpix;
    int face;
    float mult;
};

struct Lifted {
    float fu, fv;         // ERP flow of this face at this pixel
    int xi, yi;           // floor(fu), floor(fv)
    float fx, fy;         // fu - floor(fu), fv - floor(fv): bilinear fractions of the end point
};

// floor / fraction without the XU pipe (|x| < 2^22): round to nearest with the 1.5*2^23 trick, then fix down
__device__ __forceinline__ void split_floor(float x, int& xi, float& frac) {
    const float y = x + 12582912.0f;
    xi = __float_as_int(y) - 0x4B400000;
    frac = x - (y - 12582912.0f);
    if (frac < 0.0f) {
        frac += 1.0f;
        xi -= 1;
    }
}

// End point of one accepted sample, projection_icosahedron.py:336-366, written as two rotations:
//   direction of the end point  d = (gx, cos(phi0) - gy sin(phi0), sin(phi0) + gy cos(phi0))
//       [reverse_gnomonic_projection (gnomonic_projection.py:76-85) with sin c = rho cos c divided out;
//        gx, gy = pixel2gnomonic (:186-194) of (ix + u, iy + v)]
//   delta_theta  = atan2 of (d.x, d.y) rotated back by (theta_src - theta0)      -> u = delta_theta * W / 2 pi
//   delta_colat  = atan2 of (hypot(d.x, d.y), d.z) rotated back by colat_src     -> v = delta_colat * H / pi
// Both arctangents see the small angle between end point and source pixel, so float32 keeps a relative
// precision of 1e-7 on the flow itself.  delta_theta in (-pi, pi] is the reference's flow after its seam
// fix (:350-361); without wrap_around the reference leaves theta_t wrapped into [-pi, pi) instead
// (spherical_coordinates.py:237,59), restored at the end.
//
// PRECISE = false: float32 throughout (end point away from the poles).
// PRECISE = true : direction and rotations in float64, CUDA's double atan2/sqrt (rows next to the poles,
//                  where d.x, d.y cancel and a rotation by a float32 sine loses the small polar radius).
template <bool PRECISE>
__device__ __forceinline__ Lifted lift_endpoint(const SampleGeom& G, float u, float v, float sin_colat_src,
                                                float cos_colat_src, double sin_cs64, double cos_cs64, double sin_d64,
                                                double cos_d64, float theta_src, float u_scale, float v_scale,
                                                double u_scale64, double v_scale64, float erp_wf, bool wrap_around) {
    Lifted o;
    if (PRECISE) {
        const double a = fma((double)u, G.kx, G.gx0);
        const double b = fma((double)v, G.ky_sin, G.bz0);
        const double c = fma(-(double)v, G.ky_cos, G.cy0);
        const double xr = fma(b, cos_d64, a * sin_d64);
        const double yr = fma(a, cos_d64, -b * sin_d64);
        double dtheta = atan2(yr, xr);
        const double h = sqrt(fma(xr, xr, yr * yr));
        const double xc = fma(c, cos_cs64, h * sin_cs64);
        const double yc = fma(h, cos_cs64, -c * sin_cs64);
        const double dcolat = atan2(yc, xc);
        // bilinear fractions from the wrapped (small) flow; the seam convention then moves whole images
        const double fu = dtheta * u_scale64, fv = dcolat * v_scale64;
        const double fiu = floor(fu), fiv = floor(fv);
        o.xi = (int)fiu;
        o.yi = (int)fiv;
        o.fx = (float)(fu - fiu);
        o.fy = (float)(fv - fiv);
        o.fu = (float)fu;
        o.fv = (float)fv;
        if (!wrap_around) {
            const double tt = (double)theta_src + dtheta;
            if (tt >= TIF_PI) { o.fu -= erp_wf; o.xi -= (int)erp_wf; }
            else if (tt < -TIF_PI) { o.fu += erp_wf; o.xi += (int)erp_wf; }
        }
        return o;
    }
    // displacement of the end point from the source pixel on the tangent plane (gnomonic units, y up)
    const float dx = fmaf(u, G.kxf, G.d0x);
    const float dy = fmaf(-v, G.kyf, G.d0y);
    // east / north / radial components in the source pixel's local frame: small terms stay small, so
    // float32 keeps ~1e-7 relative precision on the flow itself instead of on the O(1) direction vector
    const float E = fmaf(dy, G.e1, dx * G.e2);
    const float N = fmaf(dy, G.n1, dx * G.n2);
    const float U = fmaf(dy, G.u1, fmaf(dx, G.u2, G.u0));
    // sin_colat_src = cos(phi_src), cos_colat_src = sin(phi_src)
    const float rm = fmaf(U, sin_colat_src, -N * cos_colat_src);       // horizontal part in the source meridian plane
    const float z = fmaf(U, cos_colat_src, N * sin_colat_src);         // part along the polar axis
    float dtheta = atan2_fast(E, rm);
    const float h = sqrt_newton(fmaxf(fmaf(E, E, rm * rm), 1e-30f));
    // h cos(colat_s) - z sin(colat_s) = -N + (h - rm) cos(colat_s), and h - rm = E^2 / (h + rm)
    const float yc = rm > 0.0f ? fmaf(E * E, cos_colat_src * rcp_newton(h + rm), -N)
                               : fmaf(h, cos_colat_src, -z * sin_colat_src);
    const float xc = fmaf(z, cos_colat_src, h * sin_colat_src);
    const float dcolat = atan2_fast(yc, xc);
    o.fu = dtheta * u_scale;
    o.fv = dcolat * v_scale;
    split_floor(o.fu, o.xi, o.fx);
    split_floor(o.fv, o.yi, o.fy);
    if (!wrap_around) {
        const float tt = theta_src + dtheta;
        if (tt >= (float)TIF_PI) { o.fu -= erp_wf; o.xi -= (int)erp_wf; }
        else if (tt < -(float)TIF_PI) { o.fu += erp_wf; o.xi += (int)erp_wf; }
    }
    return o;
}

// byte k of w as float via the 2^23 mantissa trick (PRMT + FADD, both full rate; I2F would use the XU pipe)
__device__ __forceinline__ float byte_f(uint32_t w, unsigned k) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | k)) - 8388608.0f;
}

// sum over channels of |bilinear(tar, end point) - src[r,c]| with scipy mode='constant', cval=255
// (projection.py:52-55): any coordinate outside [0, n-1] reads 255 outright.
__device__ __forceinline__ float warp_error_sum(const uint8_t* __restrict__ tar, bool tail_image, float s0, float s1,
                                                float s2, const Lifted& L, int r, int c, int erp_h, int erp_w) {
    int x0 = c + L.xi, y0 = r + L.yi;
    float fx = L.fx, fy = L.fy;
    float t0, t1, t2;
    // inside iff 0 <= x0 + fx <= W-1 and 0 <= y0 + fy <= H-1
    const bool x_ok = ((unsigned)x0 < (unsigned)(erp_w - 1)) | ((x0 == erp_w - 1) & (fx == 0.0f));
    const bool y_ok = ((unsigned)y0 < (unsigned)(erp_h - 1)) | ((y0 == erp_h - 1) & (fy == 0.0f));
    if (!(x_ok & y_ok)) {
        t0 = t1 = t2 = 255.0f;
    } else {
        if (x0 == erp_w - 1) { x0 = erp_w - 2; fx = 1.0f; }
        if (y0 == erp_h - 1) { y0 = erp_h - 2; fy = 1.0f; }
        const float wx0 = 1.0f - fx, wy0 = 1.0f - fy;
        const float w00 = wy0 * wx0, w01 = wy0 * fx, w10 = fy * wx0, w11 = fy * fx;
        const uint8_t* r0 = tar + (unsigned)((y0 * erp_w + x0) * 3);
        const uint8_t* r1 = r0 + (unsigned)(erp_w * 3);
        uint32_t lo0, hi0, lo1, hi1;
        if (tail_image & (y0 == erp_h - 2) & (x0 >= erp_w - 6)) {   // wide loads could leave the buffer
            load_tap_pair_bytes(r0, lo0, hi0);
            load_tap_pair_bytes(r1, lo1, hi1);
        } else {
            load_tap_pair(r0, (unsigned)(reinterpret_cast<uintptr_t>(r0) & 3), lo0, hi0);
            load_tap_pair(r1, (unsigned)(reinterpret_cast<uintptr_t>(r1) & 3), lo1, hi1);
        }
        t0 = byte_f(lo0, 0) * w00 + byte_f(lo0, 3) * w01 + byte_f(lo1, 0) * w10 + byte_f(lo1, 3) * w11;
        t1 = byte_f(lo0, 1) * w00 + byte_f(hi0, 0) * w01 + byte_f(lo1, 1) * w10 + byte_f(hi1, 0) * w11;
        t2 = byte_f(lo0, 2) * w00 + byte_f(hi0, 1) * w01 + byte_f(lo1, 2) * w10 + byte_f(hi1, 1) * w11;
    }
    return fabsf(t0 - s0) + fabsf(t1 - s1) + fabsf(t2 - s2);
}

__device__ __forceinline__ void load_face_consts(FaceConst* sh, const TifFace* __restrict__ faces, int n_faces) {
    for (int f = threadIdx.x; f < n_faces; f += blockDim.x) {
        const TifFace& F = faces[f];
        FaceConst fc;
        fc.sin_phi0 = F.sin_phi0;
        fc.cos_phi0 = F.cos_phi0;
        fc.xmin = F.xmin;
        fc.ymax = F.ymax;
        fc.kx = 1.0 / F.p2g_x;
        fc.ky = 1.0 / F.p2g_y;
        fc.ky_sin = fc.ky * F.sin_phi0;
        fc.ky_cos = fc.ky * F.cos_phi0;
        fc.pix_offset = F.pix_offset;
        fc.pad = 0;
        sh[f] = fc;
    }
}

// MODE 0: straightforward blend.
// MODE 1: normwarp pass A: lifts every accepted sample once, stores (fu, fv, sum_ch |diff|) in the sample
//         scratch and accumulates the face-global warp-error sums; k_blend_normwarp finishes the job.
// POLAR selects the float64 end-point lift; it is launched on the polar row bands only.
// The launch covers two row segments, [seg0_row, seg0_row + seg0_rows) then [seg1_row, ...), seg_rows rows in
// all; a warp owns an 8 x 4 pixel tile of them (compact footprint in the tangent flows and the target image).
template <int MODE, bool POLAR>
__global__ void __launch_bounds__(TIF_STITCH_THREADS, POLAR ? 4 : TIF_STITCH_MIN_BLOCKS)
k_stitch(const TifFace* __restrict__ faces, int n_faces, const uint8_t* __restrict__ s2_count,
         const uint32_t* __restrict__ s2_entries, const double* __restrict__ col_theta,
         const float2* __restrict__ row_sc, const float2* __restrict__ col_sc,
         const double2* __restrict__ row_sc64, const double2* __restrict__ col_sc64,
         const double* __restrict__ face_mult_sum, const float* __restrict__ tflow,
         const uint8_t* __restrict__ erp_src, const uint8_t* __restrict__ erp_tar,
         unsigned long long* __restrict__ nw_sums /* [B][F] fixed point */, const uint32_t* __restrict__ s2_offset,
         float* __restrict__ nw_samples /* [B][3][S]: fu, fv, |diff| sum of every accepted sample */, int64_t n_samples,
         float* __restrict__ out, int erp_h, int erp_w, int tangent_w, int n_tangent, int batch, int wrap_around,
         int seg0_row, int seg0_rows, int seg1_row, int seg_rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FaceConst* sh_face = reinterpret_cast<FaceConst*>(smem_raw);
    unsigned* sh_sum = reinterpret_cast<unsigned*>(sh_face + n_faces);      // MODE 1: per-block partial sums

    const int b0 = blockIdx.y * TIF_BATCH_CHUNK;
    load_face_consts(sh_face, faces, n_faces);
    if (MODE == 1) {
        for (int i = threadIdx.x; i < TIF_BATCH_CHUNK * n_faces; i += blockDim.x) sh_sum[i] = 0u;
    }
    __syncthreads();

    const int n_pix = erp_h * erp_w;
    int vrow, vcol;
    const bool live = tile_pixel((int64_t)blockIdx.x * blockDim.x + threadIdx.x, erp_w, seg_rows, vrow, vcol);
    const int pix = !live ? 0 : (vrow < seg0_rows ? seg0_row + vrow : seg1_row + (vrow - seg0_rows)) * erp_w + vcol;
    int cnt = 0, r = 0, c = 0;
    float theta_src = 0.0f, sin_cs = 0.0f, cos_cs = 1.0f;
    double sin_cs64 = 0.0, cos_cs64 = 1.0;
    uint32_t sample0 = 0;
    if (live) {
        cnt = s2_count[pix];
        if (MODE == 1) sample0 = s2_offset[pix];
        r = pix / erp_w;
        c = pix - r * erp_w;
        theta_src = (float)col_theta[c];
        const float2 rs = row_sc[r];        // (sin phi, cos phi) of the source row = (cos, sin) of its colatitude
        cos_cs = rs.x;
        sin_cs = rs.y;
    }
    const bool polar = POLAR;
    double sin_phi64 = 0.0, cos_phi64 = 1.0;
    if (live) {
        const double2 rs = row_sc64[r];
        sin_phi64 = rs.x;
        cos_phi64 = rs.y;
        cos_cs64 = rs.x;
        sin_cs64 = rs.y;
    }
    const double u_scale64 = (double)erp_w / TIF_TWO_PI, v_scale64 = (double)erp_h / TIF_PI;
    const float u_scale = (float)u_scale64, v_scale = (float)v_scale64;
    float acc_u[TIF_BATCH_CHUNK], acc_v[TIF_BATCH_CHUNK], acc_w[TIF_BATCH_CHUNK];
    float sp[TIF_BATCH_CHUNK][3];
#pragma unroll
    for (int bb = 0; bb < TIF_BATCH_CHUNK; ++bb) {
        acc_u[bb] = acc_v[bb] = acc_w[bb] = 0.0f;
        sp[bb][0] = sp[bb][1] = sp[bb][2] = 0.0f;
    }
    // addresses inside the chunk of TIF_BATCH_CHUNK images are 32-bit offsets from per-chunk base pointers
    // (4 images of flows / pixels / samples stay far below 4 GB), which keeps 64-bit integer maths out of the loop
    const unsigned image_bytes = (unsigned)n_pix * 3u;
    const float2* __restrict__ flow_chunk = reinterpret_cast<const float2*>(tflow) + (size_t)b0 * n_tangent;
    const uint8_t* __restrict__ tar_chunk = erp_tar + (size_t)b0 * image_bytes;
    float* __restrict__ smp_chunk = nw_samples + (size_t)b0 * 3 * n_samples;
    const unsigned n_smp = (unsigned)n_samples;
    if (MODE != 0 && live) {
#pragma unroll
        for (int bb = 0; bb < TIF_BATCH_CHUNK; ++bb) {
            if (b0 + bb < batch) {
                const uint8_t* q = erp_src + (size_t)b0 * image_bytes + ((unsigned)bb * image_bytes + (unsigned)pix * 3u);
                sp[bb][0] = (float)__ldg(q);
                sp[bb][1] = (float)__ldg(q + 1);
                sp[bb][2] = (float)__ldg(q + 2);
            }
        }
    }

    const int max_cnt = (MODE == 1) ? __reduce_max_sync(0xffffffffu, cnt) : cnt;
    uint32_t e_next = (cnt > 0) ? __ldg(s2_entries + pix) : 0u;   // cnt == 0 for dead lanes
    for (int k = 0; k < max_cnt; ++k) {
        const bool has = k < cnt;
        const uint32_t e = e_next;
        if (k + 1 < cnt) e_next = __ldg(s2_entries + (size_t)(k + 1) * n_pix + pix);     // prefetch the next entry
        SampleGeom G;
        double sin_d64 = 0.0, cos_d64 = 1.0;
        {
            const int f = (int)TIF_E_FACE(e);
            const FaceConst& F = sh_face[f];
            const int ix = (int)TIF_E_IX(e), iy = (int)TIF_E_IY(e);
            // pixel2gnomonic of the tangent pixel centre; the flow moves it by (u kx, -v ky)
            G.gx0 = fma((double)ix, F.kx, F.xmin);
            const double gy0 = fma(-(double)iy, F.ky, F.ymax);
            G.bz0 = fma(-gy0, F.sin_phi0, F.cos_phi0);
            G.cy0 = fma(gy0, F.cos_phi0, F.sin_phi0);
            G.kx = F.kx;
            G.ky_sin = F.ky_sin;
            G.ky_cos = F.ky_cos;
            G.kxf = (float)F.kx;
            G.kyf = (float)F.ky;
            G.tpix = F.pix_offset + iy * tangent_w + ix;
            G.face = f;
            G.mult = (float)TIF_E_MULT(e);
            double2 cs64 = make_double2(0.0, 1.0);
            if (has) cs64 = __ldg(col_sc64 + f * erp_w + c);        // (sin, cos)(theta_src - theta0), host numpy values
            sin_d64 = cs64.x;
            cos_d64 = cs64.y;
            if (!POLAR) {
                // the source pixel's own gnomonic position on this face (gnomonic_projection.py:36-44) and the
                // local east / north / radial frame there, once per (pixel, face) in float64
                const double sp = sin_phi64, cp = cos_phi64, s0 = F.sin_phi0, c0 = F.cos_phi0;
                const double cos_c = fma(c0 * cp, cos_d64, s0 * sp);
                const double inv = 1.0 / cos_c;
                const double gxs = cp * sin_d64 * inv;
                const double gys = fma(-s0 * cp, cos_d64, c0 * sp) * inv;
                G.d0x = (float)(G.gx0 - gxs);
                G.d0y = (float)(gy0 - gys);
                G.u0 = (float)inv;
                G.e1 = (float)(s0 * sin_d64);
                G.e2 = (float)cos_d64;
                G.n1 = (float)fma(s0 * sp, cos_d64, c0 * cp);
                G.n2 = (float)(-sp * sin_d64);
                G.u1 = (float)fma(-s0 * cp, cos_d64, c0 * sp);
                G.u2 = (float)(cp * sin_d64);
            }
        }
        const int f = G.face;
#pragma unroll
        for (int bb = 0; bb < TIF_BATCH_CHUNK; ++bb) {
            const int b = b0 + bb;
            if (b >= batch) break;
            float dsum = 0.0f;
            Lifted L;
            L.fu = L.fv = 0.0f;
            if (has) {
                const float2 uv = __ldg(flow_chunk + ((unsigned)bb * (unsigned)n_tangent + (unsigned)G.tpix));
                if (polar)
                    L = lift_endpoint<true>(G, uv.x, uv.y, sin_cs, cos_cs, sin_cs64, cos_cs64, sin_d64, cos_d64, theta_src,
                                            u_scale, v_scale, u_scale64, v_scale64, (float)erp_w, wrap_around != 0);
                else
                    L = lift_endpoint<false>(G, uv.x, uv.y, sin_cs, cos_cs, 0.0, 1.0, 0.0, 1.0, theta_src, u_scale, v_scale,
                                             0.0, 0.0, (float)erp_w, wrap_around != 0);
                if (MODE != 0)
                    dsum = warp_error_sum(tar_chunk + (unsigned)bb * image_bytes, b == batch - 1, sp[bb][0], sp[bb][1], sp[bb][2],
                                          L, r, c, erp_h, erp_w);
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
                if (has) {
                    float* smp = smp_chunk + ((unsigned)bb * 3u * n_smp + sample0 + (unsigned)k);
                    smp[0] = L.fu;
                    smp[n_smp] = L.fv;
                    smp[2u * n_smp] = dsum;
                }
                const unsigned fixed = has ? (unsigned)(dsum * G.mult * TIF_NW_FIXED_SCALE + 0.5f) : 0u;
                const int f0 = __shfl_sync(0xffffffffu, f, 0);
                const bool uniform = __all_sync(0xffffffffu, (!has) || f == f0) && __shfl_sync(0xffffffffu, (int)has, 0);
                if (uniform) {
                    const unsigned total = __reduce_add_sync(0xffffffffu, fixed);    // <= 32 x 765 x 7 x 1024
                    if ((threadIdx.x & 31) == 0) atomicAdd(&sh_sum[bb * n_faces + f0], total);
                } else if (has) {
                    atomicAdd(&sh_sum[bb * n_faces + f], fixed);
                }
            }
        }
    }

    if (MODE == 1) {
        __syncthreads();
        for (int i = threadIdx.x; i < TIF_BATCH_CHUNK * n_faces; i += blockDim.x) {
            const int bb = i / n_faces, f = i - bb * n_faces;
            if (b0 + bb < batch && sh_sum[i]) atomicAdd(&nw_sums[(size_t)(b0 + bb) * n_faces + f], (unsigned long long)sh_sum[i]);
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
            const float inv = 1.0f / acc_w[bb];
            u *= inv;
            v *= inv;
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

// normwarp pass B: a streaming blend of the samples pass A stored.
//   w = exp(-(mean_ch |diff|) / face-global mean)   (projection.py:57-58)
//   flow = sum_f w f / sum_f w                       (projection_icosahedron.py:384-393), then erp_of_wraparound
__global__ void __launch_bounds__(256)
k_blend_normwarp(int n_faces, const uint8_t* __restrict__ s2_count, const uint32_t* __restrict__ s2_entries,
                 const uint32_t* __restrict__ s2_offset, const double* __restrict__ face_mult_sum,
                 const unsigned long long* __restrict__ nw_sums, const float* __restrict__ nw_samples, int64_t n_samples,
                 float* __restrict__ out, int erp_h, int erp_w, int batch, int wrap_around) {
    extern __shared__ float sh_inv_mean[];       // [TIF_BATCH_CHUNK][F]  1 / (3 * face-global mean)
    const int b0 = blockIdx.y * TIF_BATCH_CHUNK;
    for (int i = threadIdx.x; i < TIF_BATCH_CHUNK * n_faces; i += blockDim.x) {
        const int bb = i / n_faces, f = i - bb * n_faces;
        float inv = 0.0f;
        if (b0 + bb < batch) {
            // np.mean over all accepted samples x 3 channels, duplicates counted (projection.py:57)
            const double total = (double)nw_sums[(size_t)(b0 + bb) * n_faces + f] / (double)TIF_NW_FIXED_SCALE;
            const double mean = total / (3.0 * face_mult_sum[f]);
            inv = (float)(1.0 / (3.0 * mean));
        }
        sh_inv_mean[i] = inv;
    }
    __syncthreads();
    const int n_pix = erp_h * erp_w;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= n_pix) return;
    const int cnt = s2_count[pix];
    const uint32_t sample0 = s2_offset[pix];
    const int c = pix % erp_w;
    float acc_u[TIF_BATCH_CHUNK], acc_v[TIF_BATCH_CHUNK], acc_w[TIF_BATCH_CHUNK];
#pragma unroll
    for (int bb = 0; bb < TIF_BATCH_CHUNK; ++bb) acc_u[bb] = acc_v[bb] = acc_w[bb] = 0.0f;
    for (int k = 0; k < cnt; ++k) {
        const int f = (int)TIF_E_FACE(__ldg(s2_entries + (size_t)k * n_pix + pix));
#pragma unroll
        for (int bb = 0; bb < TIF_BATCH_CHUNK; ++bb) {
            const int b = b0 + bb;
            if (b >= batch) break;
            const float* smp = nw_samples + (size_t)b * 3 * n_samples + sample0 + k;
            const float fu = __ldg(smp), fv = __ldg(smp + n_samples), dsum = __ldg(smp + 2 * n_samples);
            const float w = expf(-dsum * sh_inv_mean[bb * n_faces + f]);
            acc_u[bb] = fmaf(fu, w, acc_u[bb]);
            acc_v[bb] = fmaf(fv, w, acc_v[bb]);
            acc_w[bb] += w;
        }
    }
    const float half_w = 0.5f * (float)erp_w;
    const int split = (int)((double)erp_w - 1.0 - 0.5 * (double)erp_w);
#pragma unroll
    for (int bb = 0; bb < TIF_BATCH_CHUNK; ++bb) {
        const int b = b0 + bb;
        if (b >= batch) break;
        float u = acc_u[bb], v = acc_v[bb];
        if (acc_w[bb] != 0.0f) {      // projection_icosahedron.py:389-393
            const float inv = 1.0f / acc_w[bb];
            u *= inv;
            v *= inv;
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

// scratch = [B][F] uint64 face sums (padded to 256 bytes) + [B][3][S] float32 samples, S = accepted samples per pair
static size_t nw_sums_bytes(const TifPlan* plan, int batch) {
    const size_t raw = (size_t)batch * plan->n_faces * sizeof(unsigned long long);
    return (raw + 255) / 256 * 256;
}

extern "C" int64_t tif_ico2erp_scratch_bytes(const TifPlan* plan, int batch, int blend) {
    if (!plan || batch <= 0) return 0;
    if (blend != TIF_BLEND_NORMWARP) return 0;
    return (int64_t)(nw_sums_bytes(plan, batch) + (size_t)batch * 3 * (size_t)plan->accepted_unique * sizeof(float));
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
    const int threads = TIF_STITCH_THREADS;
    const int h = plan->erp_h, w = plan->erp_w;
    const int n_pix = h * w;
    const unsigned by = (unsigned)((batch + TIF_BATCH_CHUNK - 1) / TIF_BATCH_CHUNK);
    const size_t smem = sizeof(FaceConst) * plan->n_faces + sizeof(unsigned) * TIF_BATCH_CHUNK * plan->n_faces;
    unsigned long long* nw_sums = (unsigned long long*)scratch;
    float* nw_samples = scratch ? (float*)((unsigned char*)scratch + nw_sums_bytes(plan, batch)) : nullptr;
    const int64_t n_samples = plan->accepted_unique;
#define STITCH_ARGS                                                                                                 \
    plan->d_faces, plan->n_faces, plan->d_s2_count, plan->d_s2_entries, plan->d_col_theta, plan->d_row_sc,         \
        plan->d_col_sc, plan->d_row_sc64, plan->d_col_sc64, plan->d_face_mult_sum, tangent_flow, erp_src, erp_tar, \
        nw_sums, plan->d_s2_offset, nw_samples, n_samples, erp_flow, h, w, plan->tangent_w,                        \
        (int)plan->tangent_pixels, batch, wrap_around
#define STITCH_GRID(rows) dim3((unsigned)((tile_threads(w, (rows)) + threads - 1) / threads), by)
    // rows with sin(colatitude) < TIF_POLAR_SIN at either pole take the float64 lift (separate launch so
    // that its register footprint does not lower the occupancy of the bulk)
    int polar_rows = 0;
    while (polar_rows < h / 2 && sin(((double)polar_rows + 0.5) * TIF_PI / (double)h) < (double)TIF_POLAR_SIN) ++polar_rows;
    polar_rows = min((polar_rows + TIF_TILE_H - 1) / TIF_TILE_H * TIF_TILE_H, h / 2);     // whole tiles
    const int bulk_rows = h - 2 * polar_rows;
    if (blend == TIF_BLEND_STRAIGHTFORWARD) {
        if (bulk_rows > 0)
            k_stitch<0, false><<<STITCH_GRID(bulk_rows), threads, smem, stream>>>(STITCH_ARGS, polar_rows, bulk_rows, 0, bulk_rows);
        if (polar_rows > 0)
            k_stitch<0, true><<<STITCH_GRID(2 * polar_rows), threads, smem, stream>>>(STITCH_ARGS, 0, polar_rows,
                                                                                     h - polar_rows, 2 * polar_rows);
    } else {
        if (pass_mask & 1) {
            TIF_CUDA_CHECK(cudaMemsetAsync(nw_sums, 0, nw_sums_bytes(plan, batch), stream));
            if (bulk_rows > 0)
                k_stitch<1, false><<<STITCH_GRID(bulk_rows), threads, smem, stream>>>(STITCH_ARGS, polar_rows, bulk_rows, 0,
                                                                                     bulk_rows);
            if (polar_rows > 0)
                k_stitch<1, true><<<STITCH_GRID(2 * polar_rows), threads, smem, stream>>>(STITCH_ARGS, 0, polar_rows,
                                                                                         h - polar_rows, 2 * polar_rows);
            TIF_CUDA_CHECK(cudaGetLastError());
        }
        if (pass_mask & 2)
            k_blend_normwarp<<<dim3((unsigned)((n_pix + 255) / 256), by), 256,
                               sizeof(float) * TIF_BATCH_CHUNK * plan->n_faces, stream>>>(
                plan->n_faces, plan->d_s2_count, plan->d_s2_entries, plan->d_s2_offset, plan->d_face_mult_sum, nw_sums,
                nw_samples, n_samples, erp_flow, h, w, batch, wrap_around);
    }
#undef STITCH_ARGS
#undef STITCH_GRID
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
