// tif_common.cuh -- shared declarations of libtif_b200.so (sm_100a).
//
// The library has two translation units:
//   tif_plan.cu  geometry plan: fp64, compiled with -fmad=false so that every + - * / rounds once,
//                exactly like the numpy expressions of the reference it restates;
//   tif_exec.cu  per-batch execute kernels: gathers, bilinear, end-point lift, blend.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "tif_b200.h"

#define TIF_PI 3.141592653589793
#define TIF_TWO_PI 6.283185307179586
#define TIF_HALF_PI 1.5707963267948966

// stage-1 table entry flag: tangent pixel lies outside the padded triangle (full_face_image=False)
#define TIF_S1_OUTSIDE 0x80000000u

// stage-2 membership entry: face<<25 | mult<<22 | iy<<11 | ix
#define TIF_E_FACE(e) ((e) >> 25)
#define TIF_E_MULT(e) (((e) >> 22) & 7u)
#define TIF_E_IY(e) (((e) >> 11) & 2047u)
#define TIF_E_IX(e) ((e) & 2047u)

// fixed-point scale of the normwarp face-mean accumulation (integer sums are order-independent,
// which keeps the blend deterministic run to run)
// 128 threads x 765 x 7 x 1024 < 2^32: a block's partial sum fits a native 32-bit shared atomic
#define TIF_NW_FIXED_SCALE 1024.0f

// stage-1 staging: a block owns a TIF_S1_TILE_W x TIF_S1_TILE_H patch of the tangent stack; the ERP box its taps touch
// is copied asynchronously (cp.async) into a ring of TIF_S1_STAGES shared-memory stages, one image per stage.  The
// plan picks the stage size (a power of two between the bounds below) from the boxes it finds.  16 KB stages (four
// chunks per thread) measured slower at every ERP size than 8 KB ones with more patches on the direct path.
#define TIF_S1_TILE_W 16
#ifndef TIF_S1_TILE_H
#define TIF_S1_TILE_H 8
#endif
#ifndef TIF_S1_STAGES
#define TIF_S1_STAGES 4
#endif
#ifndef TIF_S1_MIN_STAGE_BYTES
#define TIF_S1_MIN_STAGE_BYTES 2048
#endif
#ifndef TIF_S1_MAX_STAGE_BYTES
#define TIF_S1_MAX_STAGE_BYTES 4096
#endif
// stage of the TMA form (one box of s1_tma_pitch[1] x s1_tma_rows bytes + 16 bytes of slack)
#ifndef TIF_S1_TMA_MAX_STAGE_BYTES
#define TIF_S1_TMA_MAX_STAGE_BYTES 4096u
#endif
// a stage size is taken when this share of the patches fits it (else the next larger one is tried)
#ifndef TIF_S1_STAGED_SHARE
#define TIF_S1_STAGED_SHARE 0.93
#endif

// normwarp: tangent pixels (work-list entries) lifted per block of k_lift_nw; their samples follow in the same block.
// A block's samples (2.9 per tangent pixel at 2K) are walked TIF_NW_BLOCK at a time, so the last step of a block is
// partly idle: 256 entries waste 4 % of the sample slots, 128 waste 18 % (1.44 -> 1.41 ms per 32 pairs).
#ifndef TIF_NW_BLOCK
#define TIF_NW_BLOCK 256
#endif
// images per thread in the stitch kernels (plan entries are read once per chunk); byte offsets into the images of a
// chunk are 32-bit, which bounds the ERP size a plan accepts
#define TIF_MAX_CHUNK_IMAGES 4

struct LiftFace {          // per-face constants of the end-point lift (k_lift), built once with the plan
    double sin_phi0, cos_phi0;
    double xmin, ymax;
    double kx, ky;              // 1 / pixel2gnomonic ratios (gnomonic_projection.py:188-194)
    int first_row;              // first row of the face in the tangent stack
    int pad[3];                 // 64 bytes: copied to shared memory in 16-byte pieces
};

struct TifPlan {
    int n_faces, erp_h, erp_w, tangent_w;
    int kind;                      // TIF_PLAN_ICOSAHEDRON | TIF_PLAN_CUBEMAP
    int64_t tangent_pixels;        // P
    int max_faces_per_pixel;       // K
    int64_t accepted_unique;
    int64_t referenced_tangent_pixels;
    int64_t device_bytes;
    TifFace* h_faces;              // host copy
    TifFace* d_faces;              // [F]
    LiftFace* d_lift_faces;        // [F]
    int32_t* d_pix_face;           // [P / Wt] face of every tangent row (ragged stacks)
    // stage 1: per tangent pixel
    uint32_t* d_s1_base;           // [P] (y0*W + x0) | TIF_S1_OUTSIDE ; taps (x0,x0+1) x (y0,y0+1) are always in range
    double2* d_s1_frac64;          // [P] the same in float64 (exact scipy weights, rounding fallback)
    double* d_s1_ex;               // [P] ERP sample x after sph2erp(modulo)   (parity export)
    double* d_s1_ey;               // [P]
    uint2* d_s1_tiles;             // [tiles_y * tiles_x] {first byte of the tile's ERP footprint in an image (16-byte
                                   //  aligned), rows | row bytes << 16}; rows == 0: footprint too large, direct gathers
    int s1_tiles_x, s1_tiles_y;
    int64_t s1_staged_tiles;
    int s1_stage_bytes;            // stage size of k_erp2ico_staged's shared-memory ring
    // TMA form of the staged resampler (k_erp2ico_tma): every patch whose ERP box fits one of two box shapes
    // (s1_tma_pitch[c] bytes x s1_tma_rows rows, c = 0 / 1) is fetched by one cp.async.bulk.tensor per image
    int s1_tma_pitch[2], s1_tma_rows;
    int64_t s1_tma_tiles;          // patches that fit (0: TMA form unused)
    // stage 2: per ERP pixel
    uint8_t* d_s2_count;           // [H*W] number of accepting faces
    uint32_t* d_s2_entries;        // [K][H*W] packed entries, ascending face order
    float2* d_s2_snap;             // [K][H*W] ERP position of the sample's tangent pixel minus the ERP pixel (x wrapped to +-W/2)
    uint8_t* d_ref_mark;           // [P] 1 where a tangent pixel is referenced by at least one ERP pixel
    int32_t* d_ref_list;           // [P_ref] the referenced tangent pixels, in 8x4-tile order (work list of k_lift), each as
                                   //  face << 25 | row inside the face << 11 | column
    double* d_face_mult_sum;       // [F] sum of multiplicities of accepted pixels (normwarp mean denominator / 3)
    // normwarp (icosahedron plans): the accepted samples regrouped by the tangent pixel they snap to
    int4* d_ref_pos;               // [P_ref] per work-list entry: own ERP position (floor(x) + 1 | floor(y) + 1 << 16, float bits of the two fractions) and its tangent pixel index
    uint2* d_smp;                  // [S] work-list order: {ERP row << 16 | column, face << 24 | plane k << 16 | multiplicity << 8 | entry % TIF_NW_BLOCK}
    uint32_t* d_blk_smp;           // [ceil(P_ref / TIF_NW_BLOCK) + 1] first sample of every block of TIF_NW_BLOCK work-list entries
    int64_t n_samples;             // S = accepted_unique
};

void tif_set_error(const char* fmt, ...);

#define TIF_CUDA_CHECK(expr)                                                                     \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            tif_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return TIF_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

// np.remainder for doubles: fmod, then move into the divisor's sign (numpy npy_divmod)
__device__ __forceinline__ double np_remainder(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0.0) != (m < 0.0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

// scipy NI_EXTEND_WRAP coordinate map: periodic with period n-1 (ni_interpolation.c map_coordinate)
__device__ __forceinline__ double scipy_wrap(double c, int n) {
    if (n <= 1) return 0.0;
    double sz = (double)(n - 1);
    if (c < 0.0) {
        c += sz * (double)((long long)(-c / sz) + 1);
    } else if (c > sz) {
        c -= sz * (double)((long long)(c / sz));
    }
    return c;
}
