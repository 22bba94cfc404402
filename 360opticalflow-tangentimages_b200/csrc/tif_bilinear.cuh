// tif_bilinear.cuh -- bilinear sampling helpers shared by the execute kernels (tif_exec.cu, tif_warp.cu):
// scipy's uint8 rounding, the aligned-word tap loads of interleaved RGB and the 2^24 fixed-point blend.
#pragma once

#include "tif_common.cuh"

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

static __device__ __noinline__ double bilinear_f64_exact(double p00, double p01, double p10, double p11, double fx, double fy) {
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
#ifdef TIF_TAP_LOAD3
    r.w2 = __ldg(q + 2);
#else
    if (o == 3) r.w2 = __ldg(q + 2);
#endif
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
// sum is within 255 * 4 * 2^-25 = 3.1e-5 (510 units of 2^-24) of t, and scipy's own float64 sum within
// 1e-13; if t + 0.5 lies within 1024 units (2^-14) of an integer the exact float64 expression decides
// instead (3.7e-4 of the pixels, i.e. about one warp in 85 takes the slow branch).
#define TIF_FIX_ONE 16777216.0
#define TIF_FIX_HALF 0x00800000u
#ifndef TIF_FIX_GUARD
#define TIF_FIX_GUARD 0x00000400u
#endif

// The accumulator carries the rounding half AND the guard: acc = sum + 2^23 + G.  The sample is undecided when the
// low 24 bits of acc are below 2 G (one AND, 2 G is a power of two); otherwise subtracting G again would not borrow
// from the top byte, so the top byte of acc itself is the rounded value.
static_assert((TIF_FIX_GUARD & (TIF_FIX_GUARD - 1u)) == 0u, "TIF_FIX_GUARD must be a power of two");
__device__ __forceinline__ uint32_t fix_bilinear(uint32_t p00, uint32_t p01, uint32_t p10, uint32_t p11, uint32_t w00,
                                                 uint32_t w01, uint32_t w10, uint32_t w11) {
    return p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + (TIF_FIX_HALF + TIF_FIX_GUARD);
}

__device__ __forceinline__ bool fix_undecided(uint32_t acc) {
    return (acc & (0x01000000u - 2u * TIF_FIX_GUARD)) == 0u;
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
