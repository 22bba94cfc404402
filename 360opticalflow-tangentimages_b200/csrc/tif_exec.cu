// tif_exec.cu -- per-batch execute kernels of libtif_b200.so (sm_100a).
//
// Everything here is a gather / stencil over HBM-resident data: no dense contraction, hence no
// tensor cores.  Geometry comes from the TifPlan (tif_plan.cu); one thread owns one output pixel
// and loops over a chunk of the batch so that plan entries are read once per chunk.
//
//   k_erp2ico_staged / k_erp2ico / k_erp2ico_tma   erp2ico_image   projection_icosahedron.py:233-245 (+ scipy map_coordinates, a18)
//   k_lift + k_gather                              ico2erp_flow    projection_icosahedron.py:333-396 (straightforward; cubemap stage)
//   k_lift_nw + k_nw_face_scale + k_blend_nw       ico2erp_flow    the same with face_blending_method="normwarp", projection.py:15-61
// (warp_backward, flow_warp_meshgrid, erp_of_wraparound: tif_warp.cu)
#include <cuda.h>           // CUtensorMap and its enums only: nothing of libcuda is linked
#include <stdlib.h>

#include "tif_common.cuh"

#ifndef TIF_BATCH_CHUNK
#define TIF_BATCH_CHUNK 4          // images per thread (plan entries amortised over the chunk)
#endif

#include "tif_bilinear.cuh"


// A warp owns an 8 x 4 tile of the tangent stack, not a 32 x 1 strip: tangent rows are tilted against ERP
// rows (steeply on the polar faces), so a strip touches ~21 cache lines per tap load and a tile ~9.
#define TIF_TILE_W 8
#define TIF_TILE_H 4

__device__ __forceinline__ bool tile_pixel(int64_t idx, int width, int rows, int& row, int& col) {
    // a plan holds fewer than 2^31 pixels of either kind: 32-bit division
    const unsigned warp = (unsigned)(idx >> 5);
    const int lane = (int)(idx & 31);
    const unsigned tiles_x = (unsigned)(width + TIF_TILE_W - 1) / TIF_TILE_W;
    const int ty = (int)(warp / tiles_x), tx = (int)(warp - (unsigned)ty * tiles_x);
    row = ty * TIF_TILE_H + lane / TIF_TILE_W;
    col = tx * TIF_TILE_W + lane % TIF_TILE_W;
    return row < rows && col < width;
}

// the same with the tile shape as a template parameter (TW x TH = 32)
template <int TW, int TH>
__device__ __forceinline__ bool tile_pixel_wh(int64_t idx, int width, int rows, int& row, int& col) {
    static_assert(TW * TH == 32, "a warp owns one tile");
    const int64_t warp = idx >> 5;
    const int lane = (int)(idx & 31);
    const int tiles_x = (width + TW - 1) / TW;
    const int ty = (int)(warp / tiles_x), tx = (int)(warp - (int64_t)ty * tiles_x);
    row = ty * TH + lane / TW;
    col = tx * TW + lane % TW;
    return row < rows && col < width;
}

template <int TW, int TH>
static int64_t tile_threads_wh(int width, int rows) {
    return (int64_t)((width + TW - 1) / TW) * ((rows + TH - 1) / TH) * 32;
}

__host__ __device__ __forceinline__ int64_t tile_threads(int width, int rows) {
    return (int64_t)((width + TIF_TILE_W - 1) / TIF_TILE_W) * ((rows + TIF_TILE_H - 1) / TIF_TILE_H) * 32;
}

#ifndef TIF_ERP2ICO_MIN_BLOCKS
#define TIF_ERP2ICO_MIN_BLOCKS 4
#endif
#ifndef TIF_ERP2ICO_THREADS
#define TIF_ERP2ICO_THREADS 256
#endif
// one tangent pixel, all images [b_begin, b_end): taps gathered straight from global memory
template <bool ALIGNED>
__device__ __forceinline__ void erp2ico_direct(int64_t p, const uint32_t* __restrict__ s1_base,
                                               const double2* __restrict__ s1_frac64, const uint8_t* __restrict__ erp,
                                               uint8_t* __restrict__ out, int64_t n_pix, int erp_h, int erp_w, int b_begin,
                                               int b_end, int full_face) {
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

template <bool ALIGNED>
__global__ void __launch_bounds__(TIF_ERP2ICO_THREADS, TIF_ERP2ICO_MIN_BLOCKS)
k_erp2ico(const uint32_t* __restrict__ s1_base, const double2* __restrict__ s1_frac64,
          const uint8_t* __restrict__ erp, uint8_t* __restrict__ out, int64_t n_pix, int tangent_w, int erp_h, int erp_w,
          int batch, int batch_per_block_row, int full_face) {
    int trow, tcol;
    if (!tile_pixel((int64_t)blockIdx.x * blockDim.x + threadIdx.x, tangent_w, (int)(n_pix / tangent_w), trow, tcol)) return;
    const int b_begin = blockIdx.y * batch_per_block_row;
    erp2ico_direct<ALIGNED>((int64_t)trow * tangent_w + tcol, s1_base, s1_frac64, erp, out, n_pix, erp_h, erp_w, b_begin,
                            min(batch, b_begin + batch_per_block_row), full_face);
}

// ---- staged variant ---------------------------------------------------------------------------------
// The direct kernel is bound by L1 wavefronts: every tap load of a warp touches 5-6 cache lines and each ERP
// byte is fetched by several warps.  Here a block owns a 16 x 8 patch of the tangent stack; the plan knows the
// bounding box of its taps (median 21 x 18 ERP pixels at 2K), and that box is brought into shared memory with
// 16-byte asynchronous copies (cp.async, no register staging), TIF_S1_STAGES - 1 images ahead of the blend.
// The taps then come from shared memory (LDS, no L1 tag traffic) and the global reads are whole aligned rows.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

#ifndef TIF_ERP2ICO_STAGED_MIN_BLOCKS
#define TIF_ERP2ICO_STAGED_MIN_BLOCKS 8
#endif
#define TIF_S1_THREADS (TIF_S1_TILE_W * TIF_S1_TILE_H)
#define TIF_S1_MAX_CHUNKS (TIF_S1_MAX_STAGE_BYTES / (16 * TIF_S1_THREADS))      // 16-byte chunks a thread copies per image
static_assert(TIF_S1_MAX_CHUNKS >= 1 && TIF_S1_MAX_CHUNKS * 16 * TIF_S1_THREADS == TIF_S1_MAX_STAGE_BYTES, "stage size must be a whole number of chunks per thread");

// one RGBA output pixel from a staged box: a0 = shared address of the aligned word holding the first tap, the
// second tap row is `pitch` bytes on, `sh` = 8 x byte phase (the same for both rows: pitch is a multiple of 16)
__device__ __forceinline__ uint32_t erp2ico_staged_pixel(uint32_t a0, uint32_t pitch, uint32_t sh, uint32_t w00,
                                                         uint32_t w01, uint32_t w10, uint32_t w11,
                                                         const double2* __restrict__ frac) {
    const uint32_t a1 = a0 + pitch;
    const uint32_t r00 = lds_u32(a0), r01 = lds_u32(a0 + 4u), r02 = lds_u32(a0 + 8u);
    const uint32_t r10 = lds_u32(a1), r11 = lds_u32(a1 + 4u), r12 = lds_u32(a1 + 8u);
    const uint32_t lo0 = __funnelshift_r(r00, r01, sh), hi0 = __funnelshift_r(r01, r02, sh);
    const uint32_t lo1 = __funnelshift_r(r10, r11, sh), hi1 = __funnelshift_r(r11, r12, sh);
    uint32_t ar = fix_bilinear(TIF_BYTE(lo0, 0), TIF_BYTE(lo0, 3), TIF_BYTE(lo1, 0), TIF_BYTE(lo1, 3), w00, w01, w10, w11);
    uint32_t ag = fix_bilinear(TIF_BYTE(lo0, 1), TIF_BYTE(hi0, 0), TIF_BYTE(lo1, 1), TIF_BYTE(hi1, 0), w00, w01, w10, w11);
    uint32_t ab = fix_bilinear(TIF_BYTE(lo0, 2), TIF_BYTE(hi0, 1), TIF_BYTE(lo1, 2), TIF_BYTE(hi1, 1), w00, w01, w10, w11);
    if (fix_undecided(ar) | fix_undecided(ag) | fix_undecided(ab)) {
        const double2 fr = *frac;       // rare: re-read instead of keeping four registers live
        ar = (uint32_t)scipy_round_u8(bilinear_f64_exact(TIF_BYTE(lo0, 0), TIF_BYTE(lo0, 3), TIF_BYTE(lo1, 0), TIF_BYTE(lo1, 3), fr.x, fr.y)) << 24;
        ag = (uint32_t)scipy_round_u8(bilinear_f64_exact(TIF_BYTE(lo0, 1), TIF_BYTE(hi0, 0), TIF_BYTE(lo1, 1), TIF_BYTE(hi1, 0), fr.x, fr.y)) << 24;
        ab = (uint32_t)scipy_round_u8(bilinear_f64_exact(TIF_BYTE(lo0, 2), TIF_BYTE(hi0, 1), TIF_BYTE(lo1, 2), TIF_BYTE(hi1, 1), fr.x, fr.y)) << 24;
    }
    return __byte_perm(__byte_perm(ar, ag, 0x0073u), ab, 0x0710u) | 0xff000000u;
}

// TIF_S1_STAGES buffers of `stage_bytes` each in dynamic shared memory; the plan picks the stage size from the boxes
// it found between TIF_S1_MIN_STAGE_BYTES and TIF_S1_MAX_STAGE_BYTES (2 KB at 1280x640, 4 KB at 2K and 4K with 480-pixel tangent images)
__global__ void __launch_bounds__(TIF_S1_THREADS, TIF_ERP2ICO_STAGED_MIN_BLOCKS)
k_erp2ico_staged(const uint32_t* __restrict__ s1_base, const double2* __restrict__ s1_frac64,
                 const uint2* __restrict__ s1_tiles, int tiles_x, uint32_t stage_bytes, const uint8_t* __restrict__ erp,
                 uint8_t* __restrict__ out, int64_t n_pix, int tangent_w, int erp_h, int erp_w, int batch,
                 int batch_per_block_row, int full_face) {
    extern __shared__ __align__(128) uint8_t stage[];
    const int tid = threadIdx.x;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int trow = ty * TIF_S1_TILE_H + tid / TIF_S1_TILE_W, tcol = tx * TIF_S1_TILE_W + tid % TIF_S1_TILE_W;
    const bool valid = trow < (int)(n_pix / tangent_w) && tcol < tangent_w;
    const int64_t p = (int64_t)trow * tangent_w + tcol;
    const int b_begin = blockIdx.y * batch_per_block_row;
    const int n_img = min(batch, b_begin + batch_per_block_row) - b_begin;
    const uint2 desc = s1_tiles[blockIdx.x];
    const uint32_t n_rows = desc.y & 0xffffu, pitch = desc.y >> 16;
    if (n_rows == 0u) {        // footprint too large for a stage: block-uniform fall back to direct gathers
        if (valid) erp2ico_direct<true>(p, s1_base, s1_frac64, erp, out, n_pix, erp_h, erp_w, b_begin, b_begin + n_img, full_face);
        return;
    }
    const uint32_t row_bytes = (uint32_t)erp_w * 3u;
    const size_t image_bytes = (size_t)erp_h * row_bytes;
    const uint32_t stage0 = smem_u32(stage);
    // the staged box is dense (rows `pitch` apart): chunk c of the box is bytes [16 c, 16 c + 16); thread t copies
    // the chunks t, t + 256, ...
    const uint32_t cpr = pitch >> 4, n_chunks = n_rows * cpr;
    uint32_t src[TIF_S1_MAX_CHUNKS];
#pragma unroll
    for (int j = 0; j < TIF_S1_MAX_CHUNKS; ++j) {
        const uint32_t c = (uint32_t)tid + (uint32_t)j * TIF_S1_THREADS, r = c / cpr;
        src[j] = c < n_chunks ? r * row_bytes + (c - r * cpr) * 16u : 0xffffffffu;
    }
    const uint32_t d0 = stage0 + (uint32_t)tid * 16u;
    const uint8_t* gnext = erp + (size_t)b_begin * image_bytes + desc.x;     // box of the next image to copy
    int inext = 0;                                                           // its index
    auto copy_image = [&](uint32_t off) {
        if (inext < n_img) {
#pragma unroll
            for (int j = 0; j < TIF_S1_MAX_CHUNKS; ++j)
                if (src[j] != 0xffffffffu) cp_async16(d0 + off + (uint32_t)j * (16u * TIF_S1_THREADS), gnext + src[j]);
        }
        cp_async_commit();          // one commit group per image, empty or not: the waits count groups
        gnext += image_bytes;
        ++inext;
    };
    constexpr int STAGES = TIF_S1_STAGES;
#pragma unroll
    for (int g = 0; g < STAGES - 1; ++g) copy_image((uint32_t)g * stage_bytes);

    uint32_t bs = 0u, w00 = 0u, w01 = 0u, w10 = 0u, w11 = 0u, tap = stage0, sh = 0u;
    bool blend = false;
    if (valid) {
        bs = s1_base[p];
        blend = !((bs & TIF_S1_OUTSIDE) && !full_face);
    }
    if (blend) {
        const double2 fr = s1_frac64[p];
        const double wx0 = 1.0 - fr.x, wx1 = 1.0 - wx0, wy0 = 1.0 - fr.y, wy1 = 1.0 - wy0;
        w00 = (uint32_t)(wy0 * wx0 * TIF_FIX_ONE + 0.5);
        w01 = (uint32_t)(wy0 * wx1 * TIF_FIX_ONE + 0.5);
        w10 = (uint32_t)(wy1 * wx0 * TIF_FIX_ONE + 0.5);
        w11 = (uint32_t)(wy1 * wx1 * TIF_FIX_ONE + 0.5);
        // byte offset of the first tap inside the staged box: rows are `pitch` apart there, `row_bytes` in the image
        const uint32_t t = (bs & 0x7fffffffu) * 3u - desc.x;
        const uint32_t dy = t / row_bytes, dxb = t - dy * row_bytes;
        const uint32_t o = dy * pitch + dxb;
        tap = stage0 + (o & ~3u);
        sh = (o & 3u) * 8u;
    }
    uint32_t* o4 = reinterpret_cast<uint32_t*>(out) + (size_t)b_begin * n_pix + p;
    // the image loop is unrolled over the stages
    for (int i0 = 0; i0 < n_img; i0 += STAGES) {
#pragma unroll
        for (int st = 0; st < STAGES; ++st) {
            if (i0 + st >= n_img) break;
            cp_async_wait<STAGES - 2>();            // this thread's chunks of image i have landed
            __syncthreads();                        // ... everyone's have, and everyone is done with image i - 1
            copy_image((uint32_t)((st + STAGES - 1) % STAGES) * stage_bytes);      // refills the buffer of image i - 1
            if (blend) *o4 = erp2ico_staged_pixel(tap + (uint32_t)st * stage_bytes, pitch, sh, w00, w01, w10, w11, s1_frac64 + p);
            else if (valid) *o4 = 0x00ffffffu;      // RGB 255, alpha 0
            o4 += n_pix;
        }
    }
    cp_async_wait<0>();
}

// ---- TMA variant -----------------------------------------------------------------------------------
// The same resampler with the box of every image brought in by ONE cp.async.bulk.tensor (TMA) issued by one thread
// and completing on an mbarrier, instead of one or two 16-byte cp.async per thread and image: the other 127 threads
// spend no instruction on the copy.  The ERP batch is a 3-D uint8 tensor [B][H][3W] (tensor map built per call on the
// host); a tensor map fixes its box shape, so the plan offers two (narrow / wide rows, one row count) and a patch
// uses the narrowest that holds its box; patches that fit neither keep the direct gathers.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Producer / consumer ring without a block-wide barrier: `full[st]` completes when the TMA box of a stage has landed
// (transaction bytes), `empty[st]` when every warp has blended from it (one arrival per warp).  Thread 0 is the only
// producer: before its own pixel of image i it refills the stage image i - 1 used, and waits for that stage's `empty`
// phase only -- the other warps have normally released it already, so the warps drift up to TIF_S1_STAGES - 1 images
// apart instead of meeting at a __syncthreads per image.
__global__ void __launch_bounds__(TIF_S1_THREADS, TIF_ERP2ICO_STAGED_MIN_BLOCKS)
k_erp2ico_tma(const __grid_constant__ CUtensorMap map_narrow, const __grid_constant__ CUtensorMap map_wide, uint32_t box_rows,
              uint32_t pitch_narrow, uint32_t pitch_wide, uint32_t stage_bytes, const uint32_t* __restrict__ s1_base,
              const double2* __restrict__ s1_frac64, const uint2* __restrict__ s1_tiles, int tiles_x,
              const uint8_t* __restrict__ erp, uint8_t* __restrict__ out, int64_t n_pix, int tangent_w, int erp_h, int erp_w,
              int batch, int batch_per_block_row, int full_face) {
    extern __shared__ __align__(128) uint8_t stage[];           // [TIF_S1_STAGES][stage_bytes], stage_bytes % 128 == 0
    __shared__ __align__(8) unsigned long long bars[2 * TIF_S1_STAGES];      // full[0..S), empty[0..S)
    constexpr int STAGES = TIF_S1_STAGES;
    const int tid = threadIdx.x;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int trow = ty * TIF_S1_TILE_H + tid / TIF_S1_TILE_W, tcol = tx * TIF_S1_TILE_W + tid % TIF_S1_TILE_W;
    const bool valid = trow < (int)(n_pix / tangent_w) && tcol < tangent_w;
    const int64_t p = (int64_t)trow * tangent_w + tcol;
    const int b_begin = blockIdx.y * batch_per_block_row;
    const int n_img = min(batch, b_begin + batch_per_block_row) - b_begin;
    const uint2 desc = s1_tiles[blockIdx.x];
    const uint32_t n_rows = desc.y & 0xffffu, need = desc.y >> 16;
    if (n_rows == 0u || n_rows > box_rows || need > pitch_wide) {      // block-uniform: no box shape holds the footprint
        if (valid) erp2ico_direct<true>(p, s1_base, s1_frac64, erp, out, n_pix, erp_h, erp_w, b_begin, b_begin + n_img, full_face);
        return;
    }
    const bool wide = need > pitch_narrow;
    const uint32_t pitch = wide ? pitch_wide : pitch_narrow;
    const CUtensorMap* map = wide ? &map_wide : &map_narrow;
    const uint32_t row_bytes = (uint32_t)erp_w * 3u;
    const uint32_t box_y = desc.x / row_bytes, box_x = desc.x - box_y * row_bytes;      // first row / first byte of the box
    const uint32_t stage0 = smem_u32(stage), full0 = smem_u32(bars), empty0 = full0 + 8u * STAGES;
    if (tid == 0) {
#pragma unroll
        for (int g = 0; g < STAGES; ++g) {
            mbar_init(full0 + 8u * g, 1u);
            mbar_init(empty0 + 8u * g, TIF_S1_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t box_bytes = pitch * box_rows;
    if (tid == 0) {
#pragma unroll
        for (int g = 0; g < STAGES - 1; ++g) {
            if (g < n_img) {
                mbar_expect_tx(full0 + 8u * g, box_bytes);
                tma_load_3d(stage0 + (uint32_t)g * stage_bytes, map, full0 + 8u * g, (int)box_x, (int)box_y, b_begin + g);
            }
        }
    }

    uint32_t bs = 0u, w00 = 0u, w01 = 0u, w10 = 0u, w11 = 0u, tap = stage0, sh = 0u;
    bool blend = false;
    if (valid) {
        bs = s1_base[p];
        blend = !((bs & TIF_S1_OUTSIDE) && !full_face);
    }
    if (blend) {
        const double2 fr = s1_frac64[p];
        const double wx0 = 1.0 - fr.x, wx1 = 1.0 - wx0, wy0 = 1.0 - fr.y, wy1 = 1.0 - wy0;
        w00 = (uint32_t)(wy0 * wx0 * TIF_FIX_ONE + 0.5);
        w01 = (uint32_t)(wy0 * wx1 * TIF_FIX_ONE + 0.5);
        w10 = (uint32_t)(wy1 * wx0 * TIF_FIX_ONE + 0.5);
        w11 = (uint32_t)(wy1 * wx1 * TIF_FIX_ONE + 0.5);
        // byte offset of the first tap inside the box: rows are `pitch` apart there, `row_bytes` in the image
        const uint32_t t = (bs & 0x7fffffffu) * 3u - desc.x;
        const uint32_t dy = t / row_bytes, dxb = t - dy * row_bytes;
        const uint32_t o = dy * pitch + dxb;
        tap = stage0 + (o & ~3u);
        sh = (o & 3u) * 8u;
    }
    uint32_t* o4 = reinterpret_cast<uint32_t*>(out) + (size_t)b_begin * n_pix + p;
    const bool lane0 = (tid & 31) == 0;
    for (int i0 = 0; i0 < n_img; i0 += STAGES) {
        const uint32_t parity = (uint32_t)(i0 / STAGES) & 1u;
#pragma unroll
        for (int st = 0; st < STAGES; ++st) {
            const int i = i0 + st;
            if (i >= n_img) break;
            if (tid == 0) {
                // producer, BEFORE its own pixel of image i: the stage image i - 1 used is refilled with image i + STAGES - 1
                // as soon as every warp has released it, so the request does not queue behind this warp's arithmetic
                const int j = i + STAGES - 1;
                if (j < n_img) {
                    const int sj = (st + STAGES - 1) % STAGES;
                    if (i >= 1) mbar_wait(empty0 + 8u * sj, st == 0 ? (parity ^ 1u) : parity);
                    mbar_expect_tx(full0 + 8u * sj, box_bytes);
                    tma_load_3d(stage0 + (uint32_t)sj * stage_bytes, map, full0 + 8u * sj, (int)box_x, (int)box_y, b_begin + j);
                }
            }
            mbar_wait(full0 + 8u * st, parity);     // the box of image i has landed
            if (blend) *o4 = erp2ico_staged_pixel(tap + (uint32_t)st * stage_bytes, pitch, sh, w00, w01, w10, w11, s1_frac64 + p);
            else if (valid) *o4 = 0x00ffffffu;      // RGB 255, alpha 0
            o4 += n_pix;
            __syncwarp();
            if (lane0) mbar_arrive(empty0 + 8u * st);           // this warp is done with stage st
        }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libtif_b200.so does not link libcuda)
typedef CUresult (*TifEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TifEncodeTiledFn tma_encoder() {
    static TifEncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (TifEncodeTiledFn)sym;
    }
    return fn;
}

// uint8 [batch][H][3W] with a box of `pitch` bytes x `rows` rows x 1 image; out-of-range parts of a box read zero
static bool tma_erp_map(TifEncodeTiledFn enc, const uint8_t* erp, int batch, int h, int w, uint32_t pitch, uint32_t rows, CUtensorMap* map) {
    const cuuint64_t dims[3] = {(cuuint64_t)w * 3u, (cuuint64_t)h, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {(cuuint64_t)w * 3u, (cuuint64_t)w * 3u * (cuuint64_t)h};
    const cuuint32_t box[3] = {pitch, rows, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(erp), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" int tif_erp2ico_u8(const TifPlan* plan, const uint8_t* erp, int batch, int full_face_image,
                              uint8_t* tangent_rgba, void* stream) {
    if (!plan || !erp || !tangent_rgba || batch <= 0) {
        tif_set_error("tif_erp2ico_u8: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    const int threads = TIF_ERP2ICO_THREADS;
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
    const bool stageable = aligned && plan->d_s1_tiles && plan->s1_staged_tiles > 0 &&
                           reinterpret_cast<uintptr_t>(erp) % 16 == 0 && image_bytes % 16 == 0 &&
                           ((size_t)plan->erp_w * 3) % 16 == 0;
    if (stageable) {
        const unsigned tiles = (unsigned)(plan->s1_tiles_x * plan->s1_tiles_y);
        int srows = 1;
        while (srows < batch && (int64_t)tiles * srows < 148 * 16) srows *= 2;
        const int sper = (batch + srows - 1) / srows;
        dim3 sgrid(tiles, (unsigned)((batch + sper - 1) / sper));
        // The TMA form is opt-in (TIF_S1_TMA=1 in the environment, read per call): measured on B200 it is 5 % slower than
        // the cp.async ring (0.44 against 0.42 ms per 32 images at 2K) -- with a block barrier per image (0.44), with the
        // producer / consumer mbarrier ring (0.45; requests issued before the producer's own pixel: 0.44), with 3 / 6 / 8
        // stages (0.46 / 0.44 / 0.50) and with a grid of 4 x 5 box shapes instead of two, which removes the over-fetch
        // of a fixed box (0.45).  Under ncu it executes as many instructions as the cp.async kernel (the copy code
        // it saves comes back as try_wait polls): the boxes arrive late.  A box is ~20 rows of ~100 bytes, and the TMA
        // unit, like the 1-D bulk copies before it, is bound by rows per clock, not by bytes.
        const char* tma_env = getenv("TIF_S1_TMA");
        TifEncodeTiledFn enc = (tma_env && tma_env[0] == '1') ? tma_encoder() : nullptr;
        CUtensorMap map_narrow, map_wide;
        const bool use_tma = enc && plan->s1_tma_tiles * 10 >= plan->s1_staged_tiles * 9 &&
                             tma_erp_map(enc, erp, batch, plan->erp_h, plan->erp_w, (uint32_t)plan->s1_tma_pitch[0], (uint32_t)plan->s1_tma_rows, &map_narrow) &&
                             tma_erp_map(enc, erp, batch, plan->erp_h, plan->erp_w, (uint32_t)plan->s1_tma_pitch[1], (uint32_t)plan->s1_tma_rows, &map_wide);
        if (use_tma) {
            // a stage holds the wide box + 16 bytes (the aligned 12-byte window of the last tap may end past its row)
            const uint32_t stage_bytes = ((uint32_t)plan->s1_tma_pitch[1] * (uint32_t)plan->s1_tma_rows + 16u + 127u) & ~127u;
            k_erp2ico_tma<<<sgrid, TIF_S1_THREADS, (size_t)TIF_S1_STAGES * stage_bytes, (cudaStream_t)stream>>>(
                map_narrow, map_wide, (uint32_t)plan->s1_tma_rows, (uint32_t)plan->s1_tma_pitch[0], (uint32_t)plan->s1_tma_pitch[1],
                stage_bytes, plan->d_s1_base, plan->d_s1_frac64, plan->d_s1_tiles, plan->s1_tiles_x, erp, tangent_rgba,
                plan->tangent_pixels, plan->tangent_w, plan->erp_h, plan->erp_w, batch, sper, full_face_image);
        } else {
            const size_t smem = (size_t)TIF_S1_STAGES * plan->s1_stage_bytes;
            k_erp2ico_staged<<<sgrid, TIF_S1_THREADS, smem, (cudaStream_t)stream>>>(
                plan->d_s1_base, plan->d_s1_frac64, plan->d_s1_tiles, plan->s1_tiles_x, (uint32_t)plan->s1_stage_bytes, erp,
                tangent_rgba, plan->tangent_pixels, plan->tangent_w, plan->erp_h, plan->erp_w, batch, sper, full_face_image);
        }
    } else if (aligned)
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

// Float images (the reference accepts any dtype: scipy samples a float image in float64 and the result is then cast
// into its integer tangent arrays, projection_icosahedron.py:231-241): one thread per tangent pixel, float64 bilinear in
// scipy's accumulation order, float64 out; the host applies the cast.  Not a hot path (the estimator feeds uint8).
template <typename T>
__global__ void __launch_bounds__(256)
k_erp2ico_float(const uint32_t* __restrict__ s1_base, const double2* __restrict__ s1_frac64, const T* __restrict__ erp,
                double* __restrict__ out, int64_t n_pix, int erp_h, int erp_w, int channels, int full_face) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pix) return;
    const int b = blockIdx.y;
    const uint32_t bs = s1_base[p];
    double* o = out + ((size_t)b * n_pix + p) * channels;
    if ((bs & TIF_S1_OUTSIDE) && !full_face) {
        for (int ch = 0; ch < channels; ++ch) o[ch] = 255.0;          // np.full(..., 255) stays where nothing is sampled
        return;
    }
    const double2 fr = s1_frac64[p];
    const T* r0 = erp + ((size_t)b * erp_h * erp_w + (size_t)(bs & 0x7fffffffu)) * channels;
    const T* r1 = r0 + (size_t)erp_w * channels;
    for (int ch = 0; ch < channels; ++ch)
        o[ch] = bilinear_f64_exact((double)__ldg(r0 + ch), (double)__ldg(r0 + channels + ch), (double)__ldg(r1 + ch),
                                   (double)__ldg(r1 + channels + ch), fr.x, fr.y);
}

extern "C" int tif_erp2ico_float(const TifPlan* plan, const void* erp, int dtype, int batch, int channels, int full_face_image,
                                 double* tangent, void* stream) {
    if (!plan || !erp || !tangent || batch <= 0 || channels <= 0 || (dtype != TIF_DTYPE_F32 && dtype != TIF_DTYPE_F64)) {
        tif_set_error("tif_erp2ico_float: invalid argument");
        return TIF_ERR_INVALID_ARG;
    }
    const dim3 grid((unsigned)((plan->tangent_pixels + 255) / 256), (unsigned)batch);
    if (dtype == TIF_DTYPE_F32)
        k_erp2ico_float<float><<<grid, 256, 0, (cudaStream_t)stream>>>(plan->d_s1_base, plan->d_s1_frac64, (const float*)erp, tangent,
                                                                      plan->tangent_pixels, plan->erp_h, plan->erp_w, channels, full_face_image);
    else
        k_erp2ico_float<double><<<grid, 256, 0, (cudaStream_t)stream>>>(plan->d_s1_base, plan->d_s1_frac64, (const double*)erp, tangent,
                                                                       plan->tangent_pixels, plan->erp_h, plan->erp_w, channels, full_face_image);
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

// ------------------------------------------------------------------------------------------------
// stage 2+3: tangent flows -> ERP flow
// ------------------------------------------------------------------------------------------------
// The end point of a flow vector depends only on the tangent pixel it starts from: several ERP pixels snap to
// the same nearest tangent pixel (5.59 M accepted samples reference 1.95 M tangent pixels at 2K), and the
// reference lifts the same end point once per sample.  Here the expensive part runs once per referenced
// tangent pixel:
//   k_lift    tangent pixel -> displacement of the flow end point from the pixel's own ERP position
//             (+ the bilinear sample of the target image at the end point, for normwarp)
//   k_gather  ERP pixel -> for every accepting face: flow = displacement + (tangent pixel position - ERP pixel),
//             seam convention, weights, blend.  The second term is a plan constant (d_s2_snap).
#ifndef TIF_STITCH_THREADS
#define TIF_STITCH_THREADS 256
#endif
// tangent pixels (or end points) whose sin(colatitude) is below this take the float64 lift: next to the poles
// the direction vector cancels in float32, and under normwarp the faces disagree by hundreds of ERP pixels
// there, so the photometric weights turn a 1e-4 px error of the sampling position into a flow error above
// the 1e-3 px tolerance
// block size and resident blocks per SM (register caps) measured best on B200 for each gather mode: 128-thread
// blocks at the same occupancy beat 256-thread ones by 3-5 % (finer-grained block turnover of short-lived threads)
#ifndef TIF_GATHER_MIN_BLOCKS
#define TIF_GATHER_MIN_BLOCKS 12
#endif
#ifndef TIF_GATHER1_MIN_BLOCKS
#define TIF_GATHER1_MIN_BLOCKS 10
#endif
#ifndef TIF_GATHER2_MIN_BLOCKS
#define TIF_GATHER2_MIN_BLOCKS 8
#endif
#ifndef TIF_GATHER2_THREADS
#define TIF_GATHER2_THREADS 128
#endif
#ifndef TIF_GATHER1_THREADS
#define TIF_GATHER1_THREADS 128
#endif
#ifndef TIF_GATHER0_THREADS
#define TIF_GATHER0_THREADS 128
#endif
#define TIF_GATHER_THREADS_OF(MODE) (((MODE) == 2 || (MODE) == 4) ? TIF_GATHER2_THREADS : ((MODE) == 1 ? TIF_GATHER1_THREADS : TIF_GATHER0_THREADS))
// images per k_gather thread: a divisor of the lift chunk (the records stay interleaved per lift chunk)
#ifndef TIF_GATHER_IMGS
#define TIF_GATHER_IMGS 4
#endif
static_assert(TIF_BATCH_CHUNK % TIF_GATHER_IMGS == 0, "k_gather groups must not straddle lift chunks");
// The float32 lift keeps ~6e-8 relative precision on the displacement.  Next to a pole a tangent-plane step turns
// into a longitude displacement of hundreds of ERP pixels and the faces disagree by as much, so under normwarp the
// photometric weights amplify the rounding of the sampling position: flow error ~ spread x position error, both
// proportional to W / sin(colatitude).  The float64 band therefore widens with the ERP width: sin(colatitude) below
// 0.1 W / 2048 (5.7 degrees at 2K, 10.8 at 3840, 11.5 at 4096), measured max error 4e-4 px on IID-noise images.
struct PolarBand {
    float sin2;       // tangent pixel: sin^2(colatitude) below this takes the float64 lift
    float tan;        // end point: horizontal part below tan x |polar part|
};

static PolarBand polar_band(int erp_w) {
    double s = 0.1 * (erp_w > 2048 ? (double)erp_w / 2048.0 : 1.0);
    if (s > 0.7) s = 0.7;
    PolarBand b;
    b.sin2 = (float)(s * s);
    b.tan = (float)(s / sqrt(1.0 - s * s));
    return b;
}
#define TIF_RGB_FIX 8192.0f           // target samples are kept as 21-bit fixed point (3 channels in 64 bits)


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

// atan(y/x) for x > 0 and |y| <= 0.4143 x: q + q s P(s); max abs error 2.3e-8
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

// atan2(y, x) in (-pi, pi].  The common case (a small angle between the flow end point and its start) takes
// the short polynomial; everything else the full octant reduction.
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

// 2^23 + byte k of w (one PRMT): differences of two such floats are the exact byte differences
__device__ __forceinline__ float byte_m(uint32_t w, unsigned k) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | k));
}

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

// byte k of w as float via the 2^23 mantissa trick (PRMT + FADD, both full rate; I2F would use the XU pipe)
__device__ __forceinline__ float byte_f(uint32_t w, unsigned k) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | k)) - 8388608.0f;
}

// float64 lift for the rare polar cases: both directions through CUDA's double atan2 / sqrt.  Everything it
// needs is rebuilt from the pixel index, so that the fast path keeps no float64 state live in registers.
// Returns the displacement in ERP pixels and, for the target sample, the absolute end point (ex + px, ey + py).
__device__ __noinline__ void lift_precise(const LiftFace* F, int tcol, int trow, float u, float v, int erp_w, int erp_h,
                                          const double* __restrict__ s1_ex, const double* __restrict__ s1_ey, int p,
                                          float* disp, double* end) {
    const double s0 = F->sin_phi0, c0 = F->cos_phi0;
    const double gx0 = fma((double)tcol, F->kx, F->xmin);
    const double gy0 = fma(-(double)(trow - F->first_row), F->ky, F->ymax);
    const double b0 = fma(-gy0, s0, c0), z0 = fma(gy0, c0, s0);
    const double gx = fma((double)u, F->kx, gx0), gy = fma(-(double)v, F->ky, gy0);
    const double b1 = fma(-gy, s0, c0), z1 = fma(gy, c0, s0);
    double d = atan2(gx, b1) - atan2(gx0, b0);
    if (d > TIF_PI) d -= TIF_TWO_PI;
    else if (d <= -TIF_PI) d += TIF_TWO_PI;
    const double dco = atan2(sqrt(fma(gx, gx, b1 * b1)), z1) - atan2(sqrt(fma(gx0, gx0, b0 * b0)), z0);
    const double px = d * ((double)erp_w / TIF_TWO_PI), py = dco * ((double)erp_h / TIF_PI);
    disp[0] = (float)px;
    disp[1] = (float)py;
    if (s1_ex) {
        end[0] = s1_ex[p] + px;
        end[1] = s1_ey[p] + py;
    }
}

// The cubemap stage's target sample (SAMPLE 2) is rounded to uint8 like scipy's output (projection.py:103-104), and its
// weight 0.95 / ||diff||_2 turns one flipped rounding into a flow error far above tolerance.  A float32 sample that
// lands within TIF_U8_ROUND_GUARD grey levels of a rounding step (the float32 lift moves the sampling position by
// ~1e-6 px, times an image gradient of up to 255 per pixel, plus the float32 blend) is therefore re-evaluated here:
// float64 end point, scipy's float64 bilinear in its accumulation order, round half up.  0.6 % of the samples.
#define TIF_U8_ROUND_GUARD 1.0e-3f
__device__ __noinline__ void sample_u8_exact(const LiftFace* F, int tcol, int trow, float u, float v, int erp_w, int erp_h,
                                             const double* __restrict__ s1_ex, const double* __restrict__ s1_ey, int p,
                                             const uint8_t* __restrict__ img, float* t) {
    float disp[2];
    double end[2];
    lift_precise(F, tcol, trow, u, v, erp_w, erp_h, s1_ex, s1_ey, p, disp, end);
    double x = end[0];
    const double y = end[1], xf = floor(x);
    if (xf >= (double)erp_w) x -= (double)erp_w;          // x is periodic: the same single fold as the float32 route
    else if (xf < 0.0) x += (double)erp_w;
    if (!(y >= 0.0 && y <= (double)(erp_h - 1) && x >= 0.0 && x <= (double)(erp_w - 1))) {
        t[0] = t[1] = t[2] = 255.0f;            // mode='constant', cval=255
        return;
    }
    const double y0f = floor(y), x0f = floor(x);
    const int y0 = (int)y0f, x0 = (int)x0f;
    const int y1 = y0 + 1 >= erp_h ? erp_h - 2 : y0 + 1, x1 = x0 + 1 >= erp_w ? erp_w - 2 : x0 + 1;      // weight 0 there
    const double fy = __dsub_rn(y, y0f), fx = __dsub_rn(x, x0f);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const double p00 = (double)__ldg(img + ((size_t)y0 * erp_w + x0) * 3 + ch), p01 = (double)__ldg(img + ((size_t)y0 * erp_w + x1) * 3 + ch);
        const double p10 = (double)__ldg(img + ((size_t)y1 * erp_w + x0) * 3 + ch), p11 = (double)__ldg(img + ((size_t)y1 * erp_w + x1) * 3 + ch);
        t[ch] = (float)scipy_round_u8(bilinear_f64_exact(p00, p01, p10, p11, fx, fy));
    }
}

// Displacement of the flow end point of ONE tangent pixel, projection_icosahedron.py:336-347:
//   pixel2gnomonic (gnomonic_projection.py:186-194) of (ix + u, iy + v), reverse_gnomonic_projection (:76-85),
//   sph_coord_modulo + sph2erp (spherical_coordinates.py:237-238, 59-60, 122-124),
// written in the local east / north / radial frame of the tangent pixel's own direction: the end point is that
// direction plus the small tangent-plane step (u kx, -v ky), so
//   delta_theta = atan2(E, rho)     delta_colat = atan2(-N + E^2 cos / (h + rho), ...)
// see small angles and float32 keeps ~1e-7 relative precision on the displacement itself.
// The result is (delta_theta W / 2 pi, delta_colat H / pi); k_gather adds the pixel's own ERP position.
// launch shape per variant, measured on B200: the sampling variants spill at 64 registers, 128 x 7 gives them 72;
// the plain variant is 5 % faster in 64-thread blocks at the same 50 % occupancy
#ifndef TIF_LIFT_MIN_BLOCKS
#define TIF_LIFT_MIN_BLOCKS 16
#endif
#ifndef TIF_LIFT_THREADS
#define TIF_LIFT_THREADS 64
#endif
#ifndef TIF_LIFT_NW_MIN_BLOCKS
#define TIF_LIFT_NW_MIN_BLOCKS 7
#endif
#ifndef TIF_LIFT_NW_THREADS
#define TIF_LIFT_NW_THREADS 128
#endif
#define TIF_LIFT_THREADS_OF(SAMPLE) ((SAMPLE) ? TIF_LIFT_NW_THREADS : TIF_LIFT_THREADS)
#define TIF_LIFT_BLOCKS_OF(SAMPLE) ((SAMPLE) ? TIF_LIFT_NW_MIN_BLOCKS : TIF_LIFT_MIN_BLOCKS)
#ifndef TIF_LIFT_GROUP
#define TIF_LIFT_GROUP 2
#endif
// images per k_lift thread: a divisor of the record chunk TIF_BATCH_CHUNK
#ifndef TIF_LIFT_IMGS
#define TIF_LIFT_IMGS 4
#endif
static_assert(TIF_BATCH_CHUNK % TIF_LIFT_IMGS == 0, "k_lift groups must not straddle record chunks");
static_assert(TIF_LIFT_IMGS % TIF_LIFT_GROUP == 0, "the images of a thread are split into whole groups");
// SAMPLE 0: displacement only.  SAMPLE 1: + float bilinear sample of the target image (normwarp: the reference
// converts the images to float64 first).  SAMPLE 2: + the sample rounded half-up to uint8 (cubemap stage: its
// images stay uint8, so scipy rounds every sample, projection.py:103-104).
template <int SAMPLE>
__global__ void __launch_bounds__(TIF_LIFT_THREADS_OF(SAMPLE), TIF_LIFT_BLOCKS_OF(SAMPLE))
k_lift(const LiftFace* __restrict__ lift_faces, int n_faces, const int32_t* __restrict__ row_face,
       const int32_t* __restrict__ ref_list, int n_ref, const double* __restrict__ s1_ex,
       const double* __restrict__ s1_ey, const float* __restrict__ tflow, const uint8_t* __restrict__ erp_tar,
       void* __restrict__ records, int erp_h, int erp_w, int tangent_w, int n_tangent, int batch, float u_scale,
       float v_scale /* W / 2 pi, H / pi: radians -> ERP pixels, rounded once on the host */, PolarBand polar) {
    constexpr bool NW = SAMPLE != 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LiftFace* sh_face = reinterpret_cast<LiftFace*>(smem_raw);
    // The thread's chain of dependent loads (work-list entry -> first row of its face -> flows) starts before the
    // shared face table is filled, so that the flows are in flight across the barrier: the kernel is latency bound.
    // the work list holds only tangent pixels that some ERP pixel reads (half of the stack), in tile order
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < n_ref;
    const uint32_t ref = live ? (uint32_t)__ldg(ref_list + idx) : 0u;       // face << 25 | row inside the face << 11 | column
    const int tcol = (int)TIF_E_IX(ref), trow = __ldg(&lift_faces[TIF_E_FACE(ref)].first_row) + (int)TIF_E_IY(ref);
    const int p = trow * tangent_w + tcol;
    const int b0 = blockIdx.y * TIF_LIFT_IMGS;
    const float2* __restrict__ flow_chunk = reinterpret_cast<const float2*>(tflow) + (size_t)b0 * n_tangent;
    float2 uvs[TIF_LIFT_IMGS];
#pragma unroll
    for (int bb = 0; bb < TIF_LIFT_IMGS; ++bb)
        uvs[bb] = (live && b0 + bb < batch) ? __ldg(flow_chunk + ((unsigned)bb * (unsigned)n_tangent + (unsigned)p)) : make_float2(0.0f, 0.0f);
    // the per-face constants come ready from the plan (the reciprocals are float64 divisions): 16-byte copies
    static_assert(sizeof(LiftFace) % 16 == 0, "LiftFace is copied in 16-byte pieces");
    for (int i = threadIdx.x; i < n_faces * (int)(sizeof(LiftFace) / 16); i += blockDim.x)
        reinterpret_cast<uint4*>(sh_face)[i] = __ldg(reinterpret_cast<const uint4*>(lift_faces) + i);
    __syncthreads();
    if (!live) return;
    const LiftFace& F = sh_face[TIF_E_FACE(ref)];

    // ---- per tangent pixel, once for the chunk of images -------------------------------------------------
    float a0, bz0, cy0;
    {
        const double gx0 = fma((double)tcol, F.kx, F.xmin);
        const double gy0 = fma(-(double)(trow - F.first_row), F.ky, F.ymax);
        // own direction (east, toward the face meridian, polar axis) before normalisation
        a0 = (float)gx0;
        bz0 = (float)fma(-gy0, F.sin_phi0, F.cos_phi0);
        cy0 = (float)fma(gy0, F.cos_phi0, F.sin_phi0);
    }
    const float hp2 = fmaf(a0, a0, bz0 * bz0), n2 = hp2 + cy0 * cy0;
    const bool polar_pixel = hp2 < polar.sin2 * n2;
    const float inv_hp = rsqrtf(fmaxf(hp2, 1e-30f)), inv_n = rsqrtf(n2);
    const float sp = cy0 * inv_n, cp = hp2 * inv_hp * inv_n;       // sin / cos of the pixel's latitude
    const float cos_d = bz0 * inv_hp, sin_d = a0 * inv_hp;        // longitude of the pixel against the face meridian
    const float s0 = (float)F.sin_phi0, c0 = (float)F.cos_phi0;
    // east / north / radial components of the tangent-plane step per pixel of flow: the pixel-to-gnomonic scales
    // (kx, -ky) are folded into the frame, so that E, N, U are plain linear forms of (u, v)
    const float kxf = (float)F.kx, nkyf = -(float)F.ky;
    const float e1 = nkyf * (s0 * sin_d), e2 = kxf * cos_d;
    const float n1 = nkyf * fmaf(s0 * sp, cos_d, c0 * cp), n2c = kxf * (-sp * sin_d);
    const float u1 = nkyf * fmaf(-s0 * cp, cos_d, c0 * sp), u2 = kxf * (cp * sin_d);
    const float u0 = n2 * inv_n;
    // own ERP position (plan stage 1), split into integer and fraction for the target sample
    int exi = 0, eyi = 0;
    float exf = 0.0f, eyf = 0.0f;
    if (NW) {
        const double ex = s1_ex[p], ey = s1_ey[p];
        const double fxi = floor(ex), fyi = floor(ey);
        exi = (int)fxi;
        eyi = (int)fyi;
        exf = (float)(ex - fxi);
        eyf = (float)(ey - fyi);
    }
    const unsigned image_bytes = (unsigned)erp_h * (unsigned)erp_w * 3u;
    const uint8_t* __restrict__ tar_chunk = NW ? erp_tar + (size_t)b0 * image_bytes : nullptr;
    // records of one tangent pixel are interleaved over the images of the chunk: k_gather fetches them from
    // one or two adjacent sectors instead of one sector per image
    const size_t rec0 = ((size_t)(b0 / TIF_BATCH_CHUNK) * n_tangent + p) * TIF_BATCH_CHUNK + (size_t)(b0 % TIF_BATCH_CHUNK);

    // The images of the chunk are handled in groups of TIF_LIFT_GROUP: first the end points of the whole group,
    // then all its target-image tap loads, then the blends -- the tap gathers are L2-latency bound, and issued
    // back to back they overlap instead of each waiting behind the lift arithmetic of its own image.
    constexpr int G = NW ? TIF_LIFT_GROUP : 1;
    constexpr unsigned kNoSample = 0xffffffffu;
#pragma unroll
    for (int g0 = 0; g0 < TIF_LIFT_IMGS; g0 += G) {
        float disp_x[G], disp_y[G], fx[G], fy[G];
        unsigned off[G];             // byte offset of the first tap in the chunk of target images
        bool tail[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int bb = g0 + j, b = b0 + bb;
            off[j] = kNoSample;
            tail[j] = false;
            disp_x[j] = disp_y[j] = fx[j] = fy[j] = 0.0f;
            if (b >= batch) continue;
            const float2 uv = uvs[bb];
            const float dx = uv.x, dy = uv.y;          // the scales live in e1 .. u2
            const float E = fmaf(dy, e1, dx * e2);
            const float N = fmaf(dy, n1, dx * n2c);
            const float U = fmaf(dy, u1, fmaf(dx, u2, u0));
            const float rm = fmaf(U, cp, -N * sp);          // horizontal part in the pixel's meridian plane
            const float z = fmaf(U, sp, N * cp);            // part along the polar axis
            const float h = sqrt_newton(fmaxf(fmaf(E, E, rm * rm), 1e-30f));
            int xi = 0, yi = 0;
            if (polar_pixel || h < polar.tan * fabsf(z)) {       // pixel or end point inside the polar band
                float disp[2];
                double end[2];
                lift_precise(&F, tcol, trow, uv.x, uv.y, erp_w, erp_h, NW ? s1_ex : nullptr, s1_ey, p, disp, end);
                disp_x[j] = disp[0];
                disp_y[j] = disp[1];
                if (NW) {
                    const double fax = floor(end[0]), fay = floor(end[1]);
                    xi = (int)fax;
                    yi = (int)fay;
                    fx[j] = (float)(end[0] - fax);
                    fy[j] = (float)(end[1] - fay);
                }
            } else {
                // atan2(E, rm) = 2 atan(E / (h + rm)) with h = |(E, rm)|: the reciprocal of h + rm is needed anyway
                // (next line), and the half angle stays inside the short polynomial up to a 45 degree longitude step
                const float r_hp = rcp_newton(fmaxf(h + rm, 1e-30f));
                const float q_half = E * r_hp;
                float dtheta;
                if (rm > 0.0f && fabsf(q_half) <= 0.4143f) {
                    const float s_half = q_half * q_half;
                    float pr = 7.901539654e-02f;
                    pr = fmaf(pr, s_half, -1.382412463e-01f);
                    pr = fmaf(pr, s_half, 1.997184753e-01f);
                    pr = fmaf(pr, s_half, -3.333275616e-01f);
                    dtheta = 2.0f * fmaf(pr, q_half * s_half, q_half);
                } else {
                    dtheta = atan2_fast(E, rm);
                }
                // h sin(lat) - z cos(lat) = -N + (h - rm) sin(lat), and h - rm = E^2 / (h + rm)
                const float yc = rm > 0.0f ? fmaf(E * E, sp * r_hp, -N) : fmaf(h, sp, -z * cp);
                const float xc = fmaf(z, sp, h * cp);
                const float dcolat = atan2_fast(yc, xc);
                disp_x[j] = dtheta * u_scale;
                disp_y[j] = dcolat * v_scale;
                if (NW) {
                    int di, dj;
                    float df, dg;
                    split_floor(disp_x[j], di, df);
                    split_floor(disp_y[j], dj, dg);
                    xi = exi + di;
                    yi = eyi + dj;
                    fx[j] = exf + df;
                    fy[j] = eyf + dg;
                    if (fx[j] >= 1.0f) { fx[j] -= 1.0f; xi += 1; }
                    if (fy[j] >= 1.0f) { fy[j] -= 1.0f; yi += 1; }
                }
            }
            if (!NW) {
                reinterpret_cast<float2*>(records)[rec0 + bb] = make_float2(disp_x[j], disp_y[j]);
                continue;
            }
            // target sample at the absolute end point, scipy mode='constant', cval=255 (projection.py:52): x is
            // periodic (the seam conventions only move the end point by whole image widths), the gap (W-1, W) and
            // anything outside [0, H-1] vertically reads 255 outright
            // own position in [-0.5, W - 0.5) plus a displacement of at most half a turn: one fold brings x into [0, W)
            if (xi >= erp_w) xi -= erp_w;
            else if (xi < 0) xi += erp_w;
            const bool x_ok = (xi < erp_w - 1) | ((xi == erp_w - 1) & (fx[j] == 0.0f));
            const bool y_ok = ((unsigned)yi < (unsigned)(erp_h - 1)) | ((yi == erp_h - 1) & (fy[j] == 0.0f));
            if (x_ok & y_ok) {
                if (xi == erp_w - 1) { xi = erp_w - 2; fx[j] = 1.0f; }
                if (yi == erp_h - 1) { yi = erp_h - 2; fy[j] = 1.0f; }
                off[j] = (unsigned)bb * image_bytes + (unsigned)((yi * erp_w + xi) * 3);
                tail[j] = (b == batch - 1) & (yi == erp_h - 2) & (xi >= erp_w - 6);     // wide loads could leave the buffer
            }
        }
        if constexpr (NW) {
        RawPair raw0[G], raw1[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            raw0[j].w0 = raw0[j].w1 = raw0[j].w2 = raw1[j].w0 = raw1[j].w1 = raw1[j].w2 = 0u;
            if (off[j] != kNoSample && !tail[j]) {
                const uint8_t* r0 = tar_chunk + off[j];
                const uint8_t* r1 = r0 + (unsigned)(erp_w * 3);
                raw0[j] = load_raw_pair(r0, (unsigned)(reinterpret_cast<uintptr_t>(r0) & 3));
                raw1[j] = load_raw_pair(r1, (unsigned)(reinterpret_cast<uintptr_t>(r1) & 3));
            }
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int bb = g0 + j;
            if (b0 + bb >= batch) continue;
            float t0 = 255.0f, t1 = 255.0f, t2 = 255.0f;
            if (off[j] != kNoSample) {
                const uint8_t* r0 = tar_chunk + off[j];
                const uint8_t* r1 = r0 + (unsigned)(erp_w * 3);
                uint32_t lo0, hi0, lo1, hi1;
                if (tail[j]) {
                    load_tap_pair_bytes(r0, lo0, hi0);
                    load_tap_pair_bytes(r1, lo1, hi1);
                } else {
                    align_pair(raw0[j], (unsigned)(reinterpret_cast<uintptr_t>(r0) & 3), lo0, hi0);
                    align_pair(raw1[j], (unsigned)(reinterpret_cast<uintptr_t>(r1) & 3), lo1, hi1);
                }
                const float wx0 = 1.0f - fx[j], wy0 = 1.0f - fy[j];
                const float w00 = wy0 * wx0, w01 = wy0 * fx[j], w10 = fy[j] * wx0, w11 = fy[j] * fx[j];
                t0 = byte_f(lo0, 0) * w00 + byte_f(lo0, 3) * w01 + byte_f(lo1, 0) * w10 + byte_f(lo1, 3) * w11;
                t1 = byte_f(lo0, 1) * w00 + byte_f(hi0, 0) * w01 + byte_f(lo1, 1) * w10 + byte_f(hi1, 0) * w11;
                t2 = byte_f(lo0, 2) * w00 + byte_f(hi0, 1) * w01 + byte_f(lo1, 2) * w10 + byte_f(hi1, 1) * w11;
            }
            if (SAMPLE == 2) {          // scipy uint8 output: trunc(t + 0.5)
                const float r0 = t0 + 0.5f, r1 = t1 + 0.5f, r2 = t2 + 0.5f;
                const float near = fminf(fminf(fabsf(r0 - rintf(r0)), fabsf(r1 - rintf(r1))), fabsf(r2 - rintf(r2)));
                if (off[j] != kNoSample && near < TIF_U8_ROUND_GUARD) {
                    float te[3];
                    sample_u8_exact(&F, tcol, trow, uvs[bb].x, uvs[bb].y, erp_w, erp_h, s1_ex, s1_ey, p,
                                    erp_tar + (size_t)(b0 + bb) * image_bytes, te);
                    t0 = te[0];
                    t1 = te[1];
                    t2 = te[2];
                } else {
                    t0 = floorf(r0);
                    t1 = floorf(r1);
                    t2 = floorf(r2);
                }
            }
            // 21-bit fixed point (x 8192, round to nearest) through the 2^23 mantissa trick: no XU conversions
            const uint32_t q0 = __float_as_uint(fmaf(t0, TIF_RGB_FIX, 8388608.0f)) - 0x4B000000u;
            const uint32_t q1 = __float_as_uint(fmaf(t1, TIF_RGB_FIX, 8388608.0f)) - 0x4B000000u;
            const uint32_t q2 = __float_as_uint(fmaf(t2, TIF_RGB_FIX, 8388608.0f)) - 0x4B000000u;
            reinterpret_cast<uint4*>(records)[rec0 + bb] =
                make_uint4(__float_as_uint(disp_x[j]), __float_as_uint(disp_y[j]), q0 | (q1 << 21), (q1 >> 11) | (q2 << 10));
        }
        }       // if constexpr (NW)
    }
}

__device__ __forceinline__ float exp2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// shared-memory reduction without a return value (and without the compiler's warp-aggregation wrapper)
__device__ __forceinline__ void red_shared_add(unsigned* addr, unsigned val) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(addr)), "r"(val) : "memory");
}

// MODE 3 / 4: cubemap2erp_flow weights (projection.py:64-117): 1 inside the unpadded square else 0 / 0.95 over the
// L2 norm of the per-channel warp error; no face-global mean, a single pass.
// MODE 0: straightforward blend.  MODE 1: normwarp pass A (face-global warp-error sums, order-independent fixed
// point).  MODE 2: normwarp pass B (weights from the sums of pass A, blend).  One thread per ERP pixel and chunk
// of images; a warp owns an 8 x 4 pixel tile.
template <int MODE>
__global__ void __launch_bounds__(TIF_GATHER_THREADS_OF(MODE), (MODE == 2 || MODE == 4) ? TIF_GATHER2_MIN_BLOCKS : (MODE == 1 ? TIF_GATHER1_MIN_BLOCKS : TIF_GATHER_MIN_BLOCKS))
k_gather(const TifFace* __restrict__ faces, int n_faces, const uint8_t* __restrict__ s2_count,
         const uint32_t* __restrict__ s2_entries, const float2* __restrict__ s2_snap,
         const double* __restrict__ face_mult_sum, const void* __restrict__ records, const uint8_t* __restrict__ erp_src,
         unsigned long long* __restrict__ nw_sums /* [B][F] fixed point */, const float* __restrict__ nw_scale /* [B][F] */,
         float* __restrict__ out, int erp_h, int erp_w, int tangent_w, int n_tangent, int batch, int wrap_around) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* sh_offset = reinterpret_cast<int*>(smem_raw);                       // [F] first pixel of each face
    unsigned* sh_sum = reinterpret_cast<unsigned*>(sh_offset + n_faces);      // MODE 1: [CHUNK][F] partial sums
    float* sh_inv_mean = reinterpret_cast<float*>(sh_offset + n_faces);      // MODE 2: [CHUNK][F] -log2(e) / (3 mean)
    const int b0 = blockIdx.y * TIF_GATHER_IMGS;         // first image of this thread's group
    const unsigned bsub = (unsigned)(b0 % TIF_BATCH_CHUNK);   // its position inside the lift chunk
    for (int f = threadIdx.x; f < n_faces; f += blockDim.x) sh_offset[f] = faces[f].pix_offset;
    if (MODE == 1) {
        for (int i = threadIdx.x; i < TIF_GATHER_IMGS * n_faces; i += blockDim.x) sh_sum[i] = 0u;
    } else if (MODE == 2) {
        // [b0 .. b0 + IMGS) x F is one contiguous run of the table k_nw_face_scale wrote
        for (int i = threadIdx.x; i < TIF_GATHER_IMGS * n_faces; i += blockDim.x)
            sh_inv_mean[i] = (b0 * n_faces + i < batch * n_faces) ? nw_scale[(size_t)b0 * n_faces + i] : 0.0f;
    }
    __syncthreads();

    const int n_pix = erp_h * erp_w;
    int r, c;
    const bool live = tile_pixel((int64_t)blockIdx.x * blockDim.x + threadIdx.x, erp_w, erp_h, r, c);
    const int pix = live ? r * erp_w + c : 0;
    const int cnt = live ? (int)s2_count[pix] : 0;
    const float wf = (float)erp_w, half_w = 0.5f * (float)erp_w, inv_w = 1.0f / (float)erp_w;
    float acc_u[TIF_GATHER_IMGS], acc_v[TIF_GATHER_IMGS], acc_w[TIF_GATHER_IMGS];
    // A float with the bits of 1024.0 and a 21-bit fixed-point sample q (x 8192) in its mantissa is 1024 + q / 8192:
    // the packed target samples become floats with one OR, and the source pixels are kept as 1024 + value, so that
    // the per-channel difference is a single exact FADD (no I2F, no rescaling multiply)
    static_assert(TIF_RGB_FIX == 8192.0f, "the exponent trick below assumes 13 fraction bits");
    constexpr uint32_t kFixExp = 0x44800000u;
    constexpr uint32_t kQ255 = 255u << 13;                                   // the hard 255 outside the image
    constexpr uint32_t kLo255 = kQ255 | (kQ255 << 21), kHi255 = (kQ255 >> 11) | (kQ255 << 10);
    float sp[TIF_GATHER_IMGS][3];          // 1024 + source pixel
#pragma unroll
    for (int bb = 0; bb < TIF_GATHER_IMGS; ++bb) {
        acc_u[bb] = acc_v[bb] = acc_w[bb] = 0.0f;
        sp[bb][0] = sp[bb][1] = sp[bb][2] = 0.0f;
    }
    const unsigned image_bytes = (unsigned)n_pix * 3u;
    constexpr bool USES_IMAGES = (MODE == 1 || MODE == 2 || MODE == 4);
    if (USES_IMAGES && live) {
#pragma unroll
        for (int bb = 0; bb < TIF_GATHER_IMGS; ++bb) {
            if (b0 + bb < batch) {
                const uint8_t* q = erp_src + (size_t)b0 * image_bytes + ((unsigned)bb * image_bytes + (unsigned)pix * 3u);
                sp[bb][0] = __uint_as_float(kFixExp + ((uint32_t)__ldg(q) << 13));
                sp[bb][1] = __uint_as_float(kFixExp + ((uint32_t)__ldg(q + 1) << 13));
                sp[bb][2] = __uint_as_float(kFixExp + ((uint32_t)__ldg(q + 2) << 13));
            }
        }
    }
    const float2* __restrict__ disp_chunk = reinterpret_cast<const float2*>(records) + (size_t)(b0 - (int)bsub) * n_tangent;
    const uint4* __restrict__ rec_chunk = reinterpret_cast<const uint4*>(records) + (size_t)(b0 - (int)bsub) * n_tangent;   // b0 = chunk * CHUNK

    const int max_cnt = (MODE == 1) ? __reduce_max_sync(0xffffffffu, cnt) : cnt;
    uint32_t e_next = 0u;
    float2 sn_next = make_float2(0.0f, 0.0f);
    if (cnt > 0) {
        e_next = __ldg(s2_entries + pix);
        sn_next = __ldg(s2_snap + pix);
    }
    for (int k = 0; k < max_cnt; ++k) {
        const bool has = k < cnt;
        const uint32_t e = e_next;
        const float2 sn = sn_next;
        if (k + 1 < cnt) {                           // next entry in flight while this one is blended
            e_next = __ldg(s2_entries + (size_t)(k + 1) * n_pix + pix);
            sn_next = __ldg(s2_snap + (size_t)(k + 1) * n_pix + pix);
        }
        const int f = (int)TIF_E_FACE(e);
        const unsigned tpix = (unsigned)(sh_offset[f] + (int)TIF_E_IY(e) * tangent_w + (int)TIF_E_IX(e));
        const float mult = (float)TIF_E_MULT(e);
        // all records of the chunk are requested before any is used (one or two adjacent sectors)
        float2 disp[TIF_GATHER_IMGS];
        uint2 rgb[TIF_GATHER_IMGS];
#pragma unroll
        for (int bb = 0; bb < TIF_GATHER_IMGS; ++bb) {
            disp[bb] = make_float2(0.0f, 0.0f);
            rgb[bb] = make_uint2(0u, 0u);
            if (has && b0 + bb < batch) {
                if (!USES_IMAGES) {
                    disp[bb] = __ldg(disp_chunk + (tpix * TIF_BATCH_CHUNK + bsub + (unsigned)bb));
                } else {
                    const uint4 rec = __ldg(rec_chunk + (tpix * TIF_BATCH_CHUNK + bsub + (unsigned)bb));
                    disp[bb] = make_float2(__uint_as_float(rec.x), __uint_as_float(rec.y));
                    rgb[bb] = make_uint2(rec.z, rec.w);
                }
            }
        }
        // MODE 1: neighbouring pixels almost always see the same face at the same k; then the warp reduces its
        // contribution first and issues one shared-memory reduction, otherwise one per lane
        int f0 = 0;
        bool uniform = false;
        if (MODE == 1) {
            f0 = __shfl_sync(0xffffffffu, f, 0);
            uniform = __all_sync(0xffffffffu, (!has) || f == f0) && __shfl_sync(0xffffffffu, (int)has, 0);
        }
#pragma unroll
        for (int bb = 0; bb < TIF_GATHER_IMGS; ++bb) {
            const int b = b0 + bb;
            if (b >= batch) break;
            float fu = disp[bb].x + sn.x;        // end point minus this ERP pixel, the short way round
            const float fv = disp[bb].y + sn.y;
            if (fu > half_w) fu -= wf;
            else if (fu < -half_w) fu += wf;
            bool x_ok = true;
            if (wrap_around) {
                // the reference samples at c + fu after its seam fix (:350-361): outside [0, W-1] reads 255
                const float tx = (float)c + fu;
                x_ok = (tx >= 0.0f) & (tx <= wf - 1.0f);
            } else {
                // the reference leaves theta_t wrapped into [-pi, pi): c + fu lies in [-0.5, W - 0.5)
                fu -= wf * floorf(((float)c + fu + 0.5f) * inv_w);
            }
            float dsum = 0.0f, dnorm2 = 0.0f;
            if (USES_IMAGES) {
                const uint32_t plo = x_ok ? rgb[bb].x : kLo255, phi = x_ok ? rgb[bb].y : kHi255;
                const uint32_t q0 = plo & 0x1fffffu, q1 = __funnelshift_r(plo, phi, 21) & 0x1fffffu, q2 = phi >> 10;
                const float d0 = __uint_as_float(kFixExp | q0) - sp[bb][0];
                const float d1 = __uint_as_float(kFixExp | q1) - sp[bb][1];
                const float d2 = __uint_as_float(kFixExp | q2) - sp[bb][2];
                dsum = fabsf(d0) + fabsf(d1) + fabsf(d2);
                dnorm2 = fmaf(d0, d0, fmaf(d1, d1, d2 * d2));
            }
            if (MODE == 0) {
                if (has) {
                    acc_u[bb] += fu;
                    acc_v[bb] += fv;
                    acc_w[bb] += 1.0f;
                }
            } else if (MODE == 1) {
                // face-global sum of |diff| (x multiplicity) as order-independent fixed point
                const unsigned fixed = has ? (unsigned)(dsum * mult * TIF_NW_FIXED_SCALE + 0.5f) : 0u;
                if (uniform) {
                    const unsigned total = __reduce_add_sync(0xffffffffu, fixed);    // <= 32 x 765 x 7 x 1024
                    if ((threadIdx.x & 31) == 0) red_shared_add(&sh_sum[bb * n_faces + f0], total);
                } else if (has) {
                    red_shared_add(&sh_sum[bb * n_faces + f], fixed);
                }
            } else if (MODE == 2) {
                if (has) {
                    // projection.py:57-58: exp(-(mean_ch |diff|) / mean_all)
                    const float w = exp2_approx(dsum * sh_inv_mean[bb * n_faces + f]);
                    acc_u[bb] = fmaf(fu, w, acc_u[bb]);
                    acc_v[bb] = fmaf(fv, w, acc_v[bb]);
                    acc_w[bb] += w;
                }
            } else {
                if (has) {
                    // MODE 3: the plan's inner-square flag (projection.py:87-93); MODE 4: 0.95 / ||diff||_2 (:109-113)
                    const float w = (MODE == 3) ? mult : (dnorm2 == 0.0f ? 1.0f : 0.95f * rsqrtf(dnorm2));
                    acc_u[bb] = fmaf(fu, w, acc_u[bb]);
                    acc_v[bb] = fmaf(fv, w, acc_v[bb]);
                    acc_w[bb] += w;
                }
            }
        }
    }

    if (MODE == 1) {
        __syncthreads();
        for (int i = threadIdx.x; i < TIF_GATHER_IMGS * n_faces; i += blockDim.x) {
            const int bb = i / n_faces, f = i - bb * n_faces;
            if (b0 + bb < batch && sh_sum[i]) atomicAdd(&nw_sums[(size_t)(b0 + bb) * n_faces + f], (unsigned long long)sh_sum[i]);
        }
        return;
    }
    if (!live) return;
    const int split = (int)((double)erp_w - 1.0 - 0.5 * (double)erp_w);
#pragma unroll
    for (int bb = 0; bb < TIF_GATHER_IMGS; ++bb) {
        const int b = b0 + bb;
        if (b >= batch) break;
        float u = acc_u[bb], v = acc_v[bb];
        if (acc_w[bb] != 0.0f) {      // projection_icosahedron.py:389-393
            // counts (modes 0, 3) and sums of weights >= 0.95 / 442 (mode 4) are normal numbers
            const float inv = rcp_newton(acc_w[bb]);
            u *= inv;
            v *= inv;
        }
        if (wrap_around) {            // flow_postproc.erp_of_wraparound, flow_postproc.py:28-42
            if (c < split) {
                if (u > half_w) u -= wf;
            } else if (c < erp_w - 1) {
                if (u < -half_w) u += wf;
            }
        }
        reinterpret_cast<float2*>(out)[(size_t)b * n_pix + pix] = make_float2(u, v);
    }
}

// ================================================================================================
// normwarp (face_blending_method="normwarp", projection.py:15-61) in two kernels
// ================================================================================================
// The weight of a sample is exp(-mean_ch|tar(end point) - src(ERP pixel)| / G_f) with G_f the mean of that warp error
// over ALL samples of the face, so every sample's error is needed twice: for G_f and for the blend.
//
//   k_lift_nw   a block lifts TIF_NW_BLOCK work-list entries (tangent pixels) x 4 images: absolute end point to HBM as
//               32-bit fixed point (32 bytes per tangent pixel and chunk), bilinear sample of the target image at the end point to
//               SHARED memory.  It then walks the plan's sample list of exactly those tangent pixels (d_smp,
//               tangent-pixel major, one sample per thread and step, coalesced): |target sample - source pixel| per
//               channel goes to the dense plane [k][pixel] of the chunk (one 16-byte store for the 4 images, read
//               back coalesced by the blend) and, times the multiplicity, as fixed point into the face sums.
//               The target samples never leave the SM, and no pass gathers them per ERP pixel.
//   k_blend_nw  one thread per ERP pixel x 4 images, faces in index order: flow = end point - ERP pixel (integers),
//               weight = 2^(error x scale_f) from the stored errors, weighted mean, final wrap.
// Measured alternatives (B200, 32 pairs of 2048x1024): the same two phases as separate kernels with the target samples
// in HBM: 0.86 + 0.88 ms against 1.54 ms fused; source pixels staged by 4-byte cp.async: 1.76 ms; the sample steps of
// a block taken three at a time with all loads up front: 1.80 ms (registers: the lift phase spills).
// launch shapes measured on B200 (32 pairs of 2048x1024): k_lift_nw 256 x 4 (64 registers) 1.34-1.37 ms; 256 x 3: 1.48-1.52,
// 128 x 7: 1.44, 128 x 6 / 128 x 8: 1.48 / 1.45-1.49, 512 x 2: 1.50.  At 64 registers the kernel spills ~100 bytes and
// its time moves by +-2 % with the register allocation (taps of 1 or 2 images grouped, multiplicity folded into the
// fixed-point scale or not); a block-uniform face flag (no per-step shuffle + vote) costs more registers than it saves.  k_blend_nw is bound by the latency of its record gathers and wants resident warps: with the
// snap offsets 64 x 20 (48 registers) 0.47 ms, 128 x 10: 0.48, 128 x 8: 0.53, 128 x 12 (spills): 0.53, 256 x 4: 0.56, gathers
// requested one sample ahead: 0.54; without them (fixed-point end points) 64 x 24 (40 registers) 0.433, 64 x 20: 0.445,
// 128 x 12: 0.438, 64 x 28 and beyond (spills): 0.52.
#ifndef TIF_NW_LIFT_MIN_BLOCKS
#define TIF_NW_LIFT_MIN_BLOCKS 4
#endif
#ifndef TIF_NW_LIFT_GROUP
#define TIF_NW_LIFT_GROUP 2
#endif
#ifndef TIF_NW_WORD_TAPS
#define TIF_NW_WORD_TAPS 1
#endif
#ifndef TIF_NW_FOLD_MULT
#define TIF_NW_FOLD_MULT 1
#endif
#ifndef TIF_NW_BLEND_THREADS
#define TIF_NW_BLEND_THREADS 64
#endif
#ifndef TIF_NW_BLEND_MIN_BLOCKS
#define TIF_NW_BLEND_MIN_BLOCKS 24
#endif
#define TIF_NW_IMGS 4
static_assert(TIF_NW_FIXED_SCALE == 1024.0f, "k_lift_nw folds the scale into the multiplicity as a shift by 10");
static_assert(TIF_NW_IMGS % TIF_NW_LIFT_GROUP == 0, "the images of a thread are split into whole groups");
static_assert(TIF_NW_IMGS == TIF_MAX_CHUNK_IMAGES, "tif_plan_create bounds the ERP size by the chunk's byte offsets");
// per-thread fixed-point error sums are flushed (warp reduction + one shared 64-bit atomic) every this many samples:
// 32 lanes x 16 samples x 765 x 7 x 1024 < 2^32
#define TIF_NW_FLUSH 16
#ifndef TIF_NW_MAX_F32_DISP
#define TIF_NW_MAX_F32_DISP 32.0f
#endif

struct LiftFrame {      // local east / north / radial frame of a tangent pixel's own direction (see k_lift)
    float sp, cp;       // sin / cos of its latitude
    float e1, e2, n1, n2c, u1, u2, u0;
    bool polar;
};

__device__ __forceinline__ LiftFrame lift_frame(const LiftFace& F, int tcol, int frow, float polar_sin2) {
    const double gx0 = fma((double)tcol, F.kx, F.xmin);
    const double gy0 = fma(-(double)frow, F.ky, F.ymax);
    // own direction (east, toward the face meridian, polar axis) before normalisation
    const float a0 = (float)gx0;
    const float bz0 = (float)fma(-gy0, F.sin_phi0, F.cos_phi0);
    const float cy0 = (float)fma(gy0, F.cos_phi0, F.sin_phi0);
    const float hp2 = fmaf(a0, a0, bz0 * bz0), n2 = hp2 + cy0 * cy0;
    LiftFrame fr;
    fr.polar = hp2 < polar_sin2 * n2;
    const float inv_hp = rsqrtf(fmaxf(hp2, 1e-30f)), inv_n = rsqrtf(n2);
    fr.sp = cy0 * inv_n;
    fr.cp = hp2 * inv_hp * inv_n;
    const float cos_d = bz0 * inv_hp, sin_d = a0 * inv_hp;        // longitude of the pixel against the face meridian
    const float s0 = (float)F.sin_phi0, c0 = (float)F.cos_phi0;
    const float kxf = (float)F.kx, nkyf = -(float)F.ky;
    fr.e1 = nkyf * (s0 * sin_d);
    fr.e2 = kxf * cos_d;
    fr.n1 = nkyf * fmaf(s0 * fr.sp, cos_d, c0 * fr.cp);
    fr.n2c = kxf * (-fr.sp * sin_d);
    fr.u1 = nkyf * fmaf(-s0 * fr.cp, cos_d, c0 * fr.sp);
    fr.u2 = kxf * (fr.cp * sin_d);
    fr.u0 = n2 * inv_n;
    return fr;
}

// float32 displacement (ERP pixels) of the end point of one flow vector; false: the end point (or the pixel) lies in
// the polar band and the float64 route has to decide
__device__ __forceinline__ bool lift_fast(const LiftFrame& fr, float2 uv, float u_scale, float v_scale, float polar_tan,
                                          float& disp_x, float& disp_y) {
    const float dx = uv.x, dy = uv.y;          // the pixel-to-gnomonic scales live in e1 .. u2
    const float E = fmaf(dy, fr.e1, dx * fr.e2);
    const float N = fmaf(dy, fr.n1, dx * fr.n2c);
    const float U = fmaf(dy, fr.u1, fmaf(dx, fr.u2, fr.u0));
    const float rm = fmaf(U, fr.cp, -N * fr.sp);          // horizontal part in the pixel's meridian plane
    const float z = fmaf(U, fr.sp, N * fr.cp);            // part along the polar axis
    const float h = sqrt_newton(fmaxf(fmaf(E, E, rm * rm), 1e-30f));
    if (fr.polar || h < polar_tan * fabsf(z)) return false;
    // atan2(E, rm) = 2 atan(E / (h + rm)) with h = |(E, rm)|: the reciprocal of h + rm is needed anyway (below), and
    // the half angle stays inside the short polynomial up to a 45 degree longitude step
    const float r_hp = rcp_newton(fmaxf(h + rm, 1e-30f));
    const float q_half = E * r_hp;
    float dtheta;
    if (rm > 0.0f && fabsf(q_half) <= 0.4143f) {
        const float s_half = q_half * q_half;
        float pr = 7.901539654e-02f;
        pr = fmaf(pr, s_half, -1.382412463e-01f);
        pr = fmaf(pr, s_half, 1.997184753e-01f);
        pr = fmaf(pr, s_half, -3.333275616e-01f);
        dtheta = 2.0f * fmaf(pr, q_half * s_half, q_half);
    } else {
        dtheta = atan2_fast(E, rm);
    }
    // h sin(lat) - z cos(lat) = -N + (h - rm) sin(lat), and h - rm = E^2 / (h + rm)
    const float yc = rm > 0.0f ? fmaf(E * E, fr.sp * r_hp, -N) : fmaf(h, fr.sp, -z * fr.cp);
    const float xc = fmaf(z, fr.sp, h * fr.cp);
    disp_x = dtheta * u_scale;
    disp_y = atan2_fast(yc, xc) * v_scale;
    // float32 keeps ~1.2e-7 of the displacement: beyond TIF_NW_MAX_F32_DISP pixels the rounding of the sampling
    // position (times the image gradient, times the disagreement of the faces, which grows with the displacement
    // itself) would show in the photometric weights, so such end points take the float64 route as well
    return fmaxf(fabsf(disp_x), fabsf(disp_y)) <= TIF_NW_MAX_F32_DISP;
}

template <bool SRC_ALIGNED /* erp_src and erp_tar are 4-byte aligned and so is every image of the batch: pixels and taps come as aligned words addressed by 32-bit word indices */>
__global__ void __launch_bounds__(TIF_NW_BLOCK, TIF_NW_LIFT_MIN_BLOCKS)
k_lift_nw(const LiftFace* __restrict__ lift_faces, int n_faces, const int32_t* __restrict__ ref_list, int n_ref,
          const int4* __restrict__ ref_pos, const double* __restrict__ s1_ex, const double* __restrict__ s1_ey,
          const uint2* __restrict__ smp, const uint32_t* __restrict__ blk_smp, const float* __restrict__ tflow,
          const uint8_t* __restrict__ erp_src, const uint8_t* __restrict__ erp_tar, float2* __restrict__ disp_rec,
          float4* __restrict__ dplane, unsigned long long* __restrict__ nw_sums, int erp_h, int erp_w, int tangent_w,
          int n_tangent, int max_k, int batch, int wrap_around, float u_scale, float v_scale, PolarBand polar, int fix_bits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // target sample of every (image, tangent pixel) of the block: value = tap (y0, x0) [exact bytes, sh_base] + a small
    // float remainder [sh_lift.xyz].  Kept apart because |sample - source pixel| is often a few grey levels while the
    // sample itself is up to 255: as one float32 its rounding (1.5e-5) would reach the weights of smooth images, whose
    // face mean is small, wherever the faces disagree by hundreds of pixels (garbage estimator flows next to a pole).
    float4* sh_lift = reinterpret_cast<float4*>(smem_raw);                                   // [4][TIF_NW_BLOCK] remainder rgb, x code
    uint32_t* sh_base = reinterpret_cast<uint32_t*>(sh_lift + TIF_NW_IMGS * TIF_NW_BLOCK);   // [4][TIF_NW_BLOCK] R | G << 8 | B << 16
    const int tid = threadIdx.x;
    const int b0 = blockIdx.y * TIF_NW_IMGS;
    const int n_img = min(TIF_NW_IMGS, batch - b0);
    const unsigned image_bytes = (unsigned)erp_h * (unsigned)erp_w * 3u;
    const unsigned row_bytes = (unsigned)erp_w * 3u;
    // a chunk past the end of the batch repeats the last image (its results land in scratch that nobody reads)
    unsigned img_off[TIF_NW_IMGS];
#pragma unroll
    for (int bb = 0; bb < TIF_NW_IMGS; ++bb) img_off[bb] = (unsigned)min(bb, n_img - 1) * image_bytes;

    // Every thread starts with a chain of dependent loads (work-list entry -> flows -> ... ; sample-list entry -> source
    // pixel -> ...), a block lives for ~13 us and 4 blocks share an SM: the start-up latency shows.  So there is no
    // barrier before the lift -- the per-face constants come straight from the (cached) global table instead of a
    // shared copy -- and the work-list record carries the tangent pixel index, so the flows are the second link of the
    // chain, not the third.
    const int idx = blockIdx.x * TIF_NW_BLOCK + tid;
    const bool live = idx < n_ref;
    uint32_t ref = 0u;
    int4 pos = make_int4(0, 0, 0, 0);
    float2 uvs[TIF_NW_IMGS];
    if (live) {
        pos = __ldg(ref_pos + idx);                  // own ERP position (floor x + 1 | floor y + 1 << 16, fractions), tangent pixel
        ref = (uint32_t)__ldg(ref_list + idx);       // face << 25 | row inside the face << 11 | column
    }
    const uint32_t s_begin = __ldg(blk_smp + blockIdx.x), s_end = __ldg(blk_smp + blockIdx.x + 1);
    const unsigned lane = (unsigned)tid & 31u;
    const uint32_t i_first = s_begin + (uint32_t)tid;
    const uint2 sm_first = i_first < s_end ? __ldg(smp + i_first) : make_uint2(0u, 0u);
    const int p = pos.w;
    if (live) {
        const float2* __restrict__ flow_chunk = reinterpret_cast<const float2*>(tflow) + (size_t)b0 * n_tangent;
#pragma unroll
        for (int bb = 0; bb < TIF_NW_IMGS; ++bb)
            uvs[bb] = __ldg(flow_chunk + ((unsigned)min(bb, n_img - 1) * (unsigned)n_tangent + (unsigned)p));
    }

    // ---- phase 1: lift one work-list entry per thread --------------------------------------------------------
    if (live) {
        const LiftFace& F = lift_faces[TIF_E_FACE(ref)];
        const int tcol = (int)TIF_E_IX(ref), frow = (int)TIF_E_IY(ref);
        const LiftFrame fr = lift_frame(F, tcol, frow, polar.sin2);
        const int pos_x = (pos.x & 0xffff) - 1, pos_y = (int)((unsigned)pos.x >> 16) - 1;
        const float exf = __int_as_float(pos.y), eyf = __int_as_float(pos.z);
        const uint8_t* __restrict__ tar_chunk = erp_tar + (size_t)b0 * image_bytes;
        float2* __restrict__ disp_out = disp_rec + ((size_t)blockIdx.y * n_tangent + p) * TIF_NW_IMGS;
        constexpr int G = TIF_NW_LIFT_GROUP;
        constexpr unsigned kNoSample = 0xffffffffu;
#pragma unroll
        for (int g0 = 0; g0 < TIF_NW_IMGS; g0 += G) {
            float fx[G], fy[G];
            int xcode[G];
            unsigned off[G];             // byte offset of the first tap in the chunk of target images
            bool tail[G];
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const int bb = g0 + j;
                const float2 uv = uvs[bb];
                float dx, dy;
                int xi, yi;
                if (lift_fast(fr, uv, u_scale, v_scale, polar.tan, dx, dy)) {
                    int di, dj;
                    float df, dg;
                    split_floor(dx, di, df);
                    split_floor(dy, dj, dg);
                    xi = pos_x + di;
                    yi = pos_y + dj;
                    fx[j] = exf + df;
                    fy[j] = eyf + dg;
                    if (fx[j] >= 1.0f) { fx[j] -= 1.0f; xi += 1; }
                    if (fy[j] >= 1.0f) { fy[j] -= 1.0f; yi += 1; }
                } else {
                    float disp[2];
                    double end[2];
                    lift_precise(&F, tcol, F.first_row + frow, uv.x, uv.y, erp_w, erp_h, s1_ex, s1_ey, p, disp, end);
                    dx = disp[0];
                    dy = disp[1];
                    const double fax = floor(end[0]), fay = floor(end[1]);
                    xi = (int)fax;
                    yi = (int)fay;
                    fx[j] = (float)(end[0] - fax);
                    fy[j] = (float)(end[1] - fay);
                }
                // target sample at the absolute end point, scipy mode='constant', cval=255 (projection.py:52): x is
                // periodic (the seam conventions only move the end point by whole image widths); own position in
                // [-0.5, W - 0.5) plus at most half a turn: one fold brings x into [0, W).  The gap (W-1, W) and
                // anything outside [0, H-1] vertically reads 255 outright.
                if (xi >= erp_w) xi -= erp_w;
                else if (xi < 0) xi += erp_w;
                // the record of the blend: the end point itself, integer part and fraction (both exact here) as one
                // fixed-point number per axis; the flow of a sample is then end point - ERP pixel in integers
                {
                    const float fix_one = (float)(1 << fix_bits);
                    const int rx = xi * (1 << fix_bits) + (__float_as_int(fmaf(fx[j], fix_one, 12582912.0f)) - 0x4B400000);
                    const int ry = yi * (1 << fix_bits) + (__float_as_int(fmaf(fy[j], fix_one, 12582912.0f)) - 0x4B400000);
                    reinterpret_cast<int2*>(disp_out)[bb] = make_int2(rx, ry);
                }
                xcode[j] = 2 * xi + (fx[j] != 0.0f ? 1 : 0);
                const bool x_ok = (xi < erp_w - 1) | ((xi == erp_w - 1) & (fx[j] == 0.0f));
                const bool y_ok = ((unsigned)yi < (unsigned)(erp_h - 1)) | ((yi == erp_h - 1) & (fy[j] == 0.0f));
                off[j] = kNoSample;
                tail[j] = false;
                if (x_ok & y_ok) {
                    if (xi == erp_w - 1) { xi = erp_w - 2; fx[j] = 1.0f; }
                    if (yi == erp_h - 1) { yi = erp_h - 2; fy[j] = 1.0f; }
                    off[j] = img_off[bb] + (unsigned)((yi * erp_w + xi) * 3);
                    // the wide loads could leave the buffer next to the end of the last image of the batch
                    tail[j] = (b0 + bb >= batch - 1) & (yi == erp_h - 2) & (xi >= erp_w - 6);
                }
            }
            RawPair raw0[G], raw1[G];
#pragma unroll
            for (int j = 0; j < G; ++j) {
                raw0[j].w0 = raw0[j].w1 = raw0[j].w2 = raw1[j].w0 = raw1[j].w1 = raw1[j].w2 = 0u;
                if (off[j] != kNoSample && !tail[j]) {
                    if (SRC_ALIGNED && TIF_NW_WORD_TAPS) {
                        // word-aligned images: 32-bit word indices from the chunk base, no pointer arithmetic
                        const uint32_t* __restrict__ q = reinterpret_cast<const uint32_t*>(tar_chunk);
                        const unsigned i0 = off[j] >> 2, i1 = (off[j] + row_bytes) >> 2;
                        raw0[j].w0 = __ldg(q + i0);
                        raw0[j].w1 = __ldg(q + i0 + 1);
                        if ((off[j] & 3u) == 3u) raw0[j].w2 = __ldg(q + i0 + 2);
                        raw1[j].w0 = __ldg(q + i1);
                        raw1[j].w1 = __ldg(q + i1 + 1);
                        if (((off[j] + row_bytes) & 3u) == 3u) raw1[j].w2 = __ldg(q + i1 + 2);
                    } else {
                        const uint8_t* r0 = tar_chunk + off[j];
                        const uint8_t* r1 = r0 + row_bytes;
                        raw0[j] = load_raw_pair(r0, (unsigned)(reinterpret_cast<uintptr_t>(r0) & 3));
                        raw1[j] = load_raw_pair(r1, (unsigned)(reinterpret_cast<uintptr_t>(r1) & 3));
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < G; ++j) {
                float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
                uint32_t base = 0x00ffffffu;                 // outside: 255 outright
                if (off[j] != kNoSample) {
                    const uint8_t* r0 = tar_chunk + off[j];
                    const uint8_t* r1 = r0 + row_bytes;
                    uint32_t lo0, hi0, lo1, hi1;
                    if (tail[j]) {
                        load_tap_pair_bytes(r0, lo0, hi0);
                        load_tap_pair_bytes(r1, lo1, hi1);
                    } else if (SRC_ALIGNED && TIF_NW_WORD_TAPS) {
                        align_pair(raw0[j], off[j] & 3u, lo0, hi0);
                        align_pair(raw1[j], (off[j] + row_bytes) & 3u, lo1, hi1);
                    } else {
                        align_pair(raw0[j], (unsigned)(reinterpret_cast<uintptr_t>(r0) & 3), lo0, hi0);
                        align_pair(raw1[j], (unsigned)(reinterpret_cast<uintptr_t>(r1) & 3), lo1, hi1);
                    }
                    // sample - tap(y0, x0) = w01 (p01 - p00) + w10 (p10 - p00) + w11 (p11 - p00): byte differences are
                    // exact (2^23 + a) - (2^23 + b), the products see small numbers
                    const float w01 = (1.0f - fy[j]) * fx[j], w10 = fy[j] * (1.0f - fx[j]), w11 = fy[j] * fx[j];
                    const float a0 = byte_m(lo0, 0), a1 = byte_m(lo0, 1), a2 = byte_m(lo0, 2);
                    t0 = fmaf(w11, byte_m(lo1, 3) - a0, fmaf(w10, byte_m(lo1, 0) - a0, w01 * (byte_m(lo0, 3) - a0)));
                    t1 = fmaf(w11, byte_m(hi1, 0) - a1, fmaf(w10, byte_m(lo1, 1) - a1, w01 * (byte_m(hi0, 0) - a1)));
                    t2 = fmaf(w11, byte_m(hi1, 1) - a2, fmaf(w10, byte_m(lo1, 2) - a2, w01 * (byte_m(hi0, 1) - a2)));
                    base = lo0 & 0x00ffffffu;
                }
                sh_lift[(g0 + j) * TIF_NW_BLOCK + tid] = make_float4(t0, t1, t2, __int_as_float(xcode[j]));
                sh_base[(g0 + j) * TIF_NW_BLOCK + tid] = base;
            }
        }
    }
    __syncthreads();

    // ---- phase 2: the samples of this block's tangent pixels, one per thread and step ------------------------
    // Software pipeline, two steps deep: the list entry of step i + 2 and the source pixels of step i + 1 are in
    // flight while step i is evaluated (the chain entry -> source pixel is two dependent global loads, and a block
    // has only about three steps to hide them in).
    const uint8_t* __restrict__ src_chunk = erp_src + (size_t)b0 * image_bytes;
    const uint32_t* __restrict__ src_words = reinterpret_cast<const uint32_t*>(src_chunk);      // SRC_ALIGNED only
    const uint32_t last_word = (uint32_t)(((size_t)n_img * image_bytes - 1) >> 2);      // last readable word of the chunk
    float4* __restrict__ dplane_chunk = dplane + (size_t)blockIdx.y * max_k * ((size_t)erp_h * erp_w);
    const unsigned n_pix = (unsigned)erp_h * (unsigned)erp_w;
    auto load_entry = [&](uint32_t i) { return i < s_end ? __ldg(smp + i) : make_uint2(0u, 0u); };
    struct SrcWords {
        uint32_t w[2 * TIF_NW_IMGS];
    };
    // source pixel of a sample in the 4 images: 3 bytes at an arbitrary phase, fetched as two aligned words each.
    // The images of a chunk are a whole number of words apart (SRC_ALIGNED), so the byte phase is the sample's own.
    auto load_src = [&](const uint2& sm, SrcWords& sw) {
        const uint32_t a = ((sm.x >> 16) * (uint32_t)erp_w + (sm.x & 0xffffu)) * 3u;
#pragma unroll
        for (int bb = 0; bb < TIF_NW_IMGS; ++bb) {
            if (SRC_ALIGNED) {
                const uint32_t wi = (a >> 2) + (img_off[bb] >> 2);
                sw.w[2 * bb] = __ldg(src_words + wi);
                // the second word is read only where it exists (the pixel's bytes never reach past the last word)
                sw.w[2 * bb + 1] = wi < last_word ? __ldg(src_words + wi + 1) : 0u;
            } else {
                const uint8_t* q = src_chunk + (img_off[bb] + a);
                sw.w[2 * bb] = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16);
                sw.w[2 * bb + 1] = 0u;
            }
        }
    };
    // Face sums: every step adds the raw bits of fma(d * mult, 1024, 2^23) = 0x4B000000 + round(d * mult * 1024) of the
    // 4 images to per-thread 32-bit accumulators (lanes without a sample have mult = 0 and add the bare offset); the
    // offsets, `pending` per lane, leave again modulo 2^32 when the warp's total is taken.  The total goes straight to
    // the global 64-bit sums as a fire-and-forget reduction (RED.ADD.64: one per warp, image and tile; integer, hence
    // order-independent): a shared-memory stage in between cost a 64-bit compare-and-swap loop per warp, a barrier and
    // a pass over the table at the end of every block -- 12 % of the kernel's stall samples.
    unsigned long long* __restrict__ sums_chunk = nw_sums + (size_t)b0 * n_faces;       // [b0 + bb][f]
    int cur_f = -1;
    unsigned pending = 0u;
    unsigned acc[TIF_NW_IMGS];
#pragma unroll
    for (int bb = 0; bb < TIF_NW_IMGS; ++bb) acc[bb] = 0u;
    auto flush = [&]() {
        if (cur_f >= 0) {
#pragma unroll
            for (int bb = 0; bb < TIF_NW_IMGS; ++bb) {
                const unsigned total = __reduce_add_sync(0xffffffffu, acc[bb] - pending * 0x4B000000u);
                if (lane == 0 && total && bb < n_img) atomicAdd(sums_chunk + bb * n_faces + cur_f, (unsigned long long)total);
                acc[bb] = 0u;
            }
        }
        pending = 0u;
    };
    uint2 sm_next = load_entry(i_first + TIF_NW_BLOCK);
    SrcWords sw_cur;
    load_src(sm_first, sw_cur);
    uint2 sm = sm_first;
    // warp-uniform trip count: the loop runs while the warp's first lane still has a sample
    for (uint32_t i = i_first; i - lane < s_end; i += TIF_NW_BLOCK) {
        const bool has = i < s_end;
        const uint2 sm_next2 = load_entry(i + 2 * TIF_NW_BLOCK);
        SrcWords sw_next;
        load_src(sm_next, sw_next);
        const int c = (int)(sm.x & 0xffffu);
        const unsigned pix = (sm.x >> 16) * (unsigned)erp_w + (unsigned)c;
        const int f = (int)(sm.y >> 24);
        const unsigned k = (sm.y >> 16) & 0xffu;
#if TIF_NW_FOLD_MULT
        const float mult_fixed = (float)(((sm.y >> 8) & 7u) << 10);          // multiplicity x TIF_NW_FIXED_SCALE; 0 for a lane without a sample
#define NW_FIXED(dv) fmaf((dv), mult_fixed, 8388608.0f)
#else
        const float mult = (float)((sm.y >> 8) & 7u);          // 0 for a lane without a sample
#define NW_FIXED(dv) fmaf((dv) * mult, TIF_NW_FIXED_SCALE, 8388608.0f)
#endif
        const unsigned local = sm.y & (TIF_NW_BLOCK - 1);
        const unsigned sh_src = ((pix * 3u) & 3u) * 8u;
        float d[TIF_NW_IMGS];
        float s0[TIF_NW_IMGS], s1[TIF_NW_IMGS], s2[TIF_NW_IMGS];
        int tmax = 0;
#pragma unroll
        for (int bb = 0; bb < TIF_NW_IMGS; ++bb) {
            const uint32_t spx = SRC_ALIGNED ? __funnelshift_r(sw_cur.w[2 * bb], sw_cur.w[2 * bb + 1], sh_src) : sw_cur.w[2 * bb];
            const float4 L = sh_lift[bb * TIF_NW_BLOCK + local];
            const uint32_t tb = sh_base[bb * TIF_NW_BLOCK + local];
            s0[bb] = byte_m(spx, 0), s1[bb] = byte_m(spx, 1), s2[bb] = byte_m(spx, 2);        // 2^23 + source pixel
            d[bb] = fabsf(L.x + (byte_m(tb, 0) - s0[bb])) + fabsf(L.y + (byte_m(tb, 1) - s1[bb])) + fabsf(L.z + (byte_m(tb, 2) - s2[bb]));
            tmax = max(tmax, abs(__float_as_int(L.w) - 2 * c));
        }
        // the reference samples at c + fu after its seam fix (:350-361): the end point taken the short way round
        // from this column; outside [0, W-1] reads 255.  2 |x_end - c| <= W <=> short way without a fold
        // (wrap_around=False leaves theta_t wrapped into [-pi, pi): always inside).  Rare: seam columns only.
        if (wrap_around && tmax > erp_w) {
#pragma unroll
            for (int bb = 0; bb < TIF_NW_IMGS; ++bb)
                if (abs(__float_as_int(sh_lift[bb * TIF_NW_BLOCK + local].w) - 2 * c) > erp_w)
                    d[bb] = (8388608.0f + 255.0f - s0[bb]) + (8388608.0f + 255.0f - s1[bb]) + (8388608.0f + 255.0f - s2[bb]);
        }
        if (has) dplane_chunk[k * n_pix + pix] = make_float4(d[0], d[1], d[2], d[3]);
        // lanes hold consecutive samples of a list sorted by tangent pixel, hence by face: the warp almost always
        // sees one face and keeps per-thread partial sums; a warp straddling two faces uses shared atomics directly
        const int f0 = __shfl_sync(0xffffffffu, f, 0);          // lane 0 is valid: the valid lanes are a prefix
        if (__all_sync(0xffffffffu, !has || f == f0)) {
            if (f0 != cur_f || pending >= TIF_NW_FLUSH) {
                flush();
                cur_f = f0;
            }
            // face-global sum of |diff| (x multiplicity) as order-independent fixed point: round to nearest through
            // the 2^23 mantissa trick (765 x 7 x 1024 < 2^23)
#pragma unroll
            for (int bb = 0; bb < TIF_NW_IMGS; ++bb) acc[bb] += __float_as_uint(NW_FIXED(d[bb]));
            ++pending;
        } else if (has) {
#pragma unroll
            for (int bb = 0; bb < TIF_NW_IMGS; ++bb)
                if (bb < n_img) atomicAdd(sums_chunk + bb * n_faces + f, (unsigned long long)(__float_as_uint(NW_FIXED(d[bb])) - 0x4B000000u));
        }
        sm = sm_next;
        sm_next = sm_next2;
        sw_cur = sw_next;
    }
    flush();
}

// grid = (ceil(W / TIF_NW_BLEND_THREADS), H, chunks): a block owns a run of one ERP row, so that row, column and pixel
// index need no division and every warp stores 256 contiguous bytes per image (what a peer-mapped output across
// NVLink wants).  Per pixel the kernel does little more than its sample loop, so the per-thread set-up counts.
__global__ void __launch_bounds__(TIF_NW_BLEND_THREADS, TIF_NW_BLEND_MIN_BLOCKS)
k_blend_nw(const TifFace* __restrict__ faces, int n_faces, const uint8_t* __restrict__ s2_count,
           const uint32_t* __restrict__ s2_entries, const int2* __restrict__ end_rec, const float4* __restrict__ dplane, const float* __restrict__ nw_scale,
           float* __restrict__ out, int erp_h, int erp_w, int tangent_w, int n_tangent, int max_k, int batch,
           int wrap_around, float inv_w, int split /* erp_of_wraparound: first column of the right half */, int fix_bits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sh_scale = reinterpret_cast<float4*>(smem_raw);                   // [F] -log2(e) / (3 mean_f) of the 4 images
    int* sh_offset = reinterpret_cast<int*>(sh_scale + n_faces);              // [F] first pixel of each face
    const int b0 = blockIdx.z * TIF_NW_IMGS;
    for (int f = threadIdx.x; f < n_faces; f += blockDim.x) {
        sh_offset[f] = faces[f].pix_offset;
        float sc[TIF_NW_IMGS];
#pragma unroll
        for (int bb = 0; bb < TIF_NW_IMGS; ++bb) sc[bb] = (b0 + bb < batch) ? nw_scale[(size_t)(b0 + bb) * n_faces + f] : 0.0f;
        sh_scale[f] = make_float4(sc[0], sc[1], sc[2], sc[3]);
    }
    const int c = blockIdx.x * TIF_NW_BLEND_THREADS + threadIdx.x;
    const unsigned n_pix = (unsigned)erp_h * (unsigned)erp_w;
    const unsigned pix = blockIdx.y * (unsigned)erp_w + (unsigned)c;
    // the first entry and error are requested before the table barrier
    const int cnt = c < erp_w ? (int)__ldg(s2_count + pix) : 0;
    const float4* __restrict__ dplane_chunk = dplane + (size_t)blockIdx.z * max_k * (size_t)n_pix + pix;
    uint32_t e_next = 0u;
    float4 d_next = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (cnt > 0) {
        e_next = __ldg(s2_entries + pix);
        d_next = __ldg(dplane_chunk);
    }
    __syncthreads();
    if (c >= erp_w) return;
    const float wf = (float)erp_w, half_w = 0.5f * wf;
    const int c_fix = c << fix_bits, r_fix = (int)blockIdx.y << fix_bits, w_fix = erp_w << fix_bits, half_fix = w_fix >> 1;
    const float fix_inv = 1.0f / (float)(1 << fix_bits);
    const float4* __restrict__ disp_chunk = reinterpret_cast<const float4*>(end_rec + (size_t)blockIdx.z * n_tangent * TIF_NW_IMGS);
    float acc_u[TIF_NW_IMGS], acc_v[TIF_NW_IMGS], acc_w[TIF_NW_IMGS];
#pragma unroll
    for (int bb = 0; bb < TIF_NW_IMGS; ++bb) acc_u[bb] = acc_v[bb] = acc_w[bb] = 0.0f;
    const uint32_t* __restrict__ e_ptr = s2_entries + pix;
    auto tangent_pixel = [&](uint32_t e) {
        return (unsigned)(sh_offset[TIF_E_FACE(e)] + (int)TIF_E_IY(e) * tangent_w + (int)TIF_E_IX(e));
    };
    for (int k = 0; k < cnt; ++k) {
        const uint32_t e = e_next;
        const float4 dd = d_next;
        if (k + 1 < cnt) {                           // next entry in flight while this one is blended
            e_ptr += n_pix;
            dplane_chunk += n_pix;
            e_next = __ldg(e_ptr);
            d_next = __ldg(dplane_chunk);
        }
        const unsigned tpix = tangent_pixel(e);
        const float4 q01 = __ldg(disp_chunk + 2u * tpix), q23 = __ldg(disp_chunk + 2u * tpix + 1u);
        const float4 sc = sh_scale[TIF_E_FACE(e)];
        const float du[TIF_NW_IMGS] = {q01.x, q01.z, q23.x, q23.z}, dv[TIF_NW_IMGS] = {q01.y, q01.w, q23.y, q23.w};
        const float de[TIF_NW_IMGS] = {dd.x, dd.y, dd.z, dd.w}, ds[TIF_NW_IMGS] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
        for (int bb = 0; bb < TIF_NW_IMGS; ++bb) {
            int fui = __float_as_int(du[bb]) - c_fix;          // end point minus this ERP pixel, the short way round
            if (fui > half_fix) fui -= w_fix;
            else if (fui < -half_fix) fui += w_fix;
            float fu = (float)fui * fix_inv;
            const float fv = (float)(__float_as_int(dv[bb]) - r_fix) * fix_inv;
            // wrap_around=False: the reference leaves theta_t wrapped into [-pi, pi): c + fu lies in [-0.5, W - 0.5)
            if (!wrap_around) fu -= wf * floorf(((float)c + fu + 0.5f) * inv_w);
            const float w = exp2_approx(de[bb] * ds[bb]);         // projection.py:57-58: exp(-(mean_ch |diff|) / mean_all)
            acc_u[bb] = fmaf(fu, w, acc_u[bb]);
            acc_v[bb] = fmaf(fv, w, acc_v[bb]);
            acc_w[bb] += w;
        }
    }
    float2* __restrict__ o = reinterpret_cast<float2*>(out) + (size_t)b0 * n_pix + pix;
#pragma unroll
    for (int bb = 0; bb < TIF_NW_IMGS; ++bb) {
        if (b0 + bb >= batch) break;
        float u = acc_u[bb], v = acc_v[bb];
        if (acc_w[bb] != 0.0f) {      // projection_icosahedron.py:389-393; the weights are flushed to zero below 2^-126, so
            const float inv = rcp_newton(acc_w[bb]);      // a non-zero sum is a normal number
            u *= inv;
            v *= inv;
        }
        if (wrap_around) {            // flow_postproc.erp_of_wraparound, flow_postproc.py:28-42
            if (c < split) {
                if (u > half_w) u -= wf;
            } else if (c < erp_w - 1) {
                if (u < -half_w) u += wf;
            }
        }
        o[(size_t)bb * n_pix] = make_float2(u, v);
    }
}

// scratch = [B][F] uint64 face sums (padded to 256 bytes) + one record per tangent pixel and image:
// float2 displacement (straightforward) or uint4 displacement + packed target sample (normwarp)
static size_t nw_sums_bytes(const TifPlan* plan, int batch) {
    const size_t raw = (size_t)batch * plan->n_faces * sizeof(unsigned long long);
    return (raw + 255) / 256 * 256;
}

// ... followed by [B][F] float: -log2(e) / (3 mean_f) of every face and image (k_nw_face_scale)
static size_t nw_header_bytes(const TifPlan* plan, int batch) {
    const size_t raw = (size_t)batch * plan->n_faces * sizeof(float);
    return nw_sums_bytes(plan, batch) + (raw + 255) / 256 * 256;
}

// The normwarp weights are exp(-sum_ch|diff| / (3 mean_f)) = 2^(sum_ch|diff| * scale_f): one thread per face and image
// turns the fixed-point sums of k_gather<1> into scale_f, instead of every block of k_gather<2> repeating the two
// float64 divisions for all its faces.
__global__ void k_nw_face_scale(const unsigned long long* __restrict__ nw_sums, const double* __restrict__ face_mult_sum,
                                int n_faces, int batch, float* __restrict__ scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch * n_faces) return;
    // np.mean over all accepted samples x 3 channels, duplicates counted (projection.py:57)
    const double total = (double)nw_sums[i] / (double)TIF_NW_FIXED_SCALE;
    const double mean = total / (3.0 * face_mult_sum[i % n_faces]);
    scale[i] = (float)(-1.4426950408889634 / (3.0 * mean));         // exp(-x / (3 mean)) = 2^(x scale)
}

// normwarp scratch after the header: displacement of every tangent pixel (float2 x 4 images per chunk) + the dense
// warp-error planes [chunk][k][H*W] (float4 = the 4 images of the chunk)
static size_t nw_disp_bytes(const TifPlan* plan, int batch) {
    const size_t chunks = (size_t)(batch + TIF_NW_IMGS - 1) / TIF_NW_IMGS;
    return chunks * (size_t)plan->tangent_pixels * TIF_NW_IMGS * sizeof(float2);
}

static size_t nw_plane_bytes(const TifPlan* plan, int batch) {
    const size_t chunks = (size_t)(batch + TIF_NW_IMGS - 1) / TIF_NW_IMGS;
    return chunks * (size_t)plan->max_faces_per_pixel * (size_t)plan->erp_h * plan->erp_w * sizeof(float4);
}

extern "C" int64_t tif_ico2erp_scratch_bytes(const TifPlan* plan, int batch, int blend) {
    if (!plan || batch <= 0) return 0;
    if (blend == TIF_BLEND_NORMWARP)
        return (int64_t)(nw_header_bytes(plan, batch) + nw_disp_bytes(plan, batch) + nw_plane_bytes(plan, batch));
    const size_t rec = (blend == TIF_BLEND_CUBEMAP_WARP_ERROR) ? sizeof(uint4) : sizeof(float2);
    const size_t padded = (size_t)(batch + TIF_BATCH_CHUNK - 1) / TIF_BATCH_CHUNK * TIF_BATCH_CHUNK;      // whole chunks
    return (int64_t)(nw_header_bytes(plan, batch) + padded * (size_t)plan->tangent_pixels * rec);
}

extern "C" int tif_ico2erp_flow_f32_ex(const TifPlan* plan, const float* tangent_flow, const uint8_t* erp_src,
                                       const uint8_t* erp_tar, int batch, int blend, int wrap_around, float* erp_flow,
                                       void* scratch, int pass_mask, void* stream_v) {
    if (!plan || !tangent_flow || !erp_flow || !scratch || batch <= 0) {
        tif_set_error("tif_ico2erp_flow_f32: invalid argument (tangent_flow, erp_flow and scratch are required)");
        return TIF_ERR_INVALID_ARG;
    }
    if (blend < TIF_BLEND_STRAIGHTFORWARD || blend > TIF_BLEND_CUBEMAP_WARP_ERROR) {
        tif_set_error("tif_ico2erp_flow_f32: unknown blend %d", blend);
        return TIF_ERR_INVALID_ARG;
    }
    const bool uses_images = blend == TIF_BLEND_NORMWARP || blend == TIF_BLEND_CUBEMAP_WARP_ERROR;
    if (uses_images && (!erp_src || !erp_tar)) {
        tif_set_error("tif_ico2erp_flow_f32: this blend needs erp_src and erp_tar");
        return TIF_ERR_INVALID_ARG;
    }
    cudaStream_t stream = (cudaStream_t)stream_v;
    const int h = plan->erp_h, w = plan->erp_w, n_faces = plan->n_faces;
    const int n_tangent = (int)plan->tangent_pixels;
    unsigned long long* nw_sums = (unsigned long long*)scratch;
    float* nw_scale = (float*)((unsigned char*)scratch + nw_sums_bytes(plan, batch));
    void* records = (unsigned char*)scratch + nw_header_bytes(plan, batch);
    const int n_ref = (int)plan->referenced_tangent_pixels;
    const unsigned lift_by = (unsigned)((batch + TIF_LIFT_IMGS - 1) / TIF_LIFT_IMGS);
    const dim3 lift_grid((unsigned)((n_ref + TIF_LIFT_THREADS - 1) / TIF_LIFT_THREADS), lift_by);
    const dim3 lift_nw_grid((unsigned)((n_ref + TIF_LIFT_NW_THREADS - 1) / TIF_LIFT_NW_THREADS), lift_by);
    const unsigned gather_by = (unsigned)((batch + TIF_GATHER_IMGS - 1) / TIF_GATHER_IMGS);
    const dim3 gather0_grid((unsigned)((tile_threads(w, h) + TIF_GATHER0_THREADS - 1) / TIF_GATHER0_THREADS), gather_by);
    const dim3 gather1_grid((unsigned)((tile_threads(w, h) + TIF_GATHER1_THREADS - 1) / TIF_GATHER1_THREADS), gather_by);
    const dim3 gather2_grid((unsigned)((tile_threads(w, h) + TIF_GATHER2_THREADS - 1) / TIF_GATHER2_THREADS), gather_by);
    const size_t lift_smem = sizeof(LiftFace) * n_faces;
    const size_t gather_smem = sizeof(int) * n_faces + sizeof(unsigned) * TIF_GATHER_IMGS * n_faces;
#define GATHER_ARGS                                                                                                  \
    plan->d_faces, n_faces, plan->d_s2_count, plan->d_s2_entries, plan->d_s2_snap, plan->d_face_mult_sum, records,   \
        erp_src, nw_sums, nw_scale, erp_flow, h, w, plan->tangent_w, n_tangent, batch, wrap_around
    const float lift_u_scale = (float)((double)w / TIF_TWO_PI), lift_v_scale = (float)((double)h / TIF_PI);
    const PolarBand polar = polar_band(w);
    const bool skip_lift = (pass_mask & 4) != 0;       // profiling: reuse the records of an earlier call
    if (blend == TIF_BLEND_STRAIGHTFORWARD || blend == TIF_BLEND_CUBEMAP_INSIDE) {
        if (!skip_lift)
            k_lift<0><<<lift_grid, TIF_LIFT_THREADS, lift_smem, stream>>>(plan->d_lift_faces, n_faces, plan->d_pix_face, plan->d_ref_list,
                                                                 n_ref, plan->d_s1_ex, plan->d_s1_ey, tangent_flow, nullptr,
                                                                 records, h, w, plan->tangent_w, n_tangent, batch, lift_u_scale, lift_v_scale, polar);
        TIF_CUDA_CHECK(cudaGetLastError());
        if (pass_mask & 3) {
            if (blend == TIF_BLEND_STRAIGHTFORWARD)
                k_gather<0><<<gather0_grid, TIF_GATHER0_THREADS, gather_smem, stream>>>(GATHER_ARGS);
            else
                k_gather<3><<<gather0_grid, TIF_GATHER0_THREADS, gather_smem, stream>>>(GATHER_ARGS);
        }
    } else if (blend == TIF_BLEND_CUBEMAP_WARP_ERROR) {
        if (!skip_lift)
            k_lift<2><<<lift_nw_grid, TIF_LIFT_NW_THREADS, lift_smem, stream>>>(plan->d_lift_faces, n_faces, plan->d_pix_face, plan->d_ref_list,
                                                             n_ref, plan->d_s1_ex, plan->d_s1_ey, tangent_flow, erp_tar,
                                                             records, h, w, plan->tangent_w, n_tangent, batch, lift_u_scale, lift_v_scale, polar);
        TIF_CUDA_CHECK(cudaGetLastError());
        if (pass_mask & 3) k_gather<4><<<gather2_grid, TIF_GATHER2_THREADS, gather_smem, stream>>>(GATHER_ARGS);
    } else {
        if (plan->kind != TIF_PLAN_ICOSAHEDRON || !plan->d_smp) {
            tif_set_error("tif_ico2erp_flow_f32: the normwarp blend needs an icosahedron plan");
            return TIF_ERR_INVALID_ARG;
        }
        const unsigned chunks = (unsigned)((batch + TIF_NW_IMGS - 1) / TIF_NW_IMGS);
        // fraction bits of the fixed-point end points in the records: 18 (4e-6 px) while (W + 2) << bits fits 31 bits
        int nw_fix_bits = 18;
        while (((long long)(w + 2) << nw_fix_bits) >= (1ll << 31)) --nw_fix_bits;
        float2* disp_rec = (float2*)records;
        float4* dplane = (float4*)((unsigned char*)records + nw_disp_bytes(plan, batch));
        if ((pass_mask & (1 | 8)) && !skip_lift) {
            TIF_CUDA_CHECK(cudaMemsetAsync(nw_sums, 0, nw_sums_bytes(plan, batch), stream));
            const dim3 grid((unsigned)((n_ref + TIF_NW_BLOCK - 1) / TIF_NW_BLOCK), chunks);
            const size_t smem = (sizeof(float4) + sizeof(uint32_t)) * TIF_NW_IMGS * TIF_NW_BLOCK;
#define LIFT_NW_ARGS                                                                                                    \
    plan->d_lift_faces, n_faces, plan->d_ref_list, n_ref, plan->d_ref_pos, plan->d_s1_ex, plan->d_s1_ey, plan->d_smp,   \
        plan->d_blk_smp, tangent_flow, erp_src, erp_tar, disp_rec, dplane, nw_sums, h, w, plan->tangent_w, n_tangent,  \
        plan->max_faces_per_pixel, batch, wrap_around, lift_u_scale, lift_v_scale, polar, nw_fix_bits
            // every image of the batch starts on a word boundary when the buffer does and H * W * 3 is a multiple of 4
            if (reinterpret_cast<uintptr_t>(erp_src) % 4 == 0 && reinterpret_cast<uintptr_t>(erp_tar) % 4 == 0 && ((size_t)h * w * 3) % 4 == 0)
                k_lift_nw<true><<<grid, TIF_NW_BLOCK, smem, stream>>>(LIFT_NW_ARGS);
            else
                k_lift_nw<false><<<grid, TIF_NW_BLOCK, smem, stream>>>(LIFT_NW_ARGS);
#undef LIFT_NW_ARGS
            TIF_CUDA_CHECK(cudaGetLastError());
        }
        if (pass_mask & 2) {
            k_nw_face_scale<<<(unsigned)((batch * n_faces + 127) / 128), 128, 0, stream>>>(nw_sums, plan->d_face_mult_sum, n_faces,
                                                                                          batch, nw_scale);
            TIF_CUDA_CHECK(cudaGetLastError());
            const dim3 grid((unsigned)((w + TIF_NW_BLEND_THREADS - 1) / TIF_NW_BLEND_THREADS), (unsigned)h, chunks);
            const size_t smem = (sizeof(float4) + sizeof(int)) * n_faces;
            k_blend_nw<<<grid, TIF_NW_BLEND_THREADS, smem, stream>>>(plan->d_faces, n_faces, plan->d_s2_count, plan->d_s2_entries,
                                                                    reinterpret_cast<const int2*>(disp_rec), dplane, nw_scale, erp_flow, h, w,
                                                                    plan->tangent_w, n_tangent, plan->max_faces_per_pixel, batch,
                                                                    wrap_around, 1.0f / (float)w,
                                                                    (int)((double)w - 1.0 - 0.5 * (double)w), nw_fix_bits);
        }
    }
#undef GATHER_ARGS
    TIF_CUDA_CHECK(cudaGetLastError());
    return TIF_OK;
}

extern "C" int tif_ico2erp_flow_f32(const TifPlan* plan, const float* tangent_flow, const uint8_t* erp_src,
                                    const uint8_t* erp_tar, int batch, int blend, int wrap_around, float* erp_flow,
                                    void* scratch, void* stream) {
    return tif_ico2erp_flow_f32_ex(plan, tangent_flow, erp_src, erp_tar, batch, blend, wrap_around, erp_flow, scratch, 3,
                                   stream);
}
