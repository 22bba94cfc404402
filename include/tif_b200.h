/*
 * tif_b200.h -- C ABI of the B200-native tangent-image hot path (libtif_b200.so).
 *
 * The reference (yuanmingze/360OpticalFlow-TangentImages) is pure Python and has no FFI / plugin
 * boundary of its own: the drop-in boundary is the Python call surface of
 * src/projection_icosahedron.py, src/flow_warp.py and src/flow_postproc.py.  Every entry point
 * below names the reference function it replaces (file:line relative to /root/reference/src); the
 * Python modules in 360opticalflow-tangentimages_b200/ keep those functions' names and argument
 * meaning and call these symbols through ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types in any signature.
 *   - All data pointers are DEVICE pointers unless the parameter is documented as host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Execute calls are
 *     asynchronous on that stream and never synchronise the host; plan creation does synchronise
 *     (it sizes its tables from a device-side count).
 *   - Caller allocates every output; nothing is retained except inside a TifPlan.
 *   - Return value: TIF_OK (0) or a negative TifStatus.  tif_status_string() names it,
 *     tif_last_error() gives the thread's last detailed message.
 *   - There is no CPU fallback anywhere: without a CUDA device every call fails with TIF_ERR_CUDA.
 *
 * Layouts
 *   ERP image        uint8  [B, H, W, 3]      interleaved RGB, W == 2H
 *   tangent stack    uint8  [B, P, 4]         RGBA; P = sum_f Ht_f * Wt; face f starts at pixel
 *                                             faces[f].pix_offset and is row-major [Ht_f, Wt]
 *                                             (for the 20-face icosahedron: [B, 20, Ht, Wt, 4])
 *   tangent flows    float32 [B, P, 2]        (u, v) in tangent pixels, same face order
 *   ERP flow         float32 [B, H, W, 2]     (u, v) in ERP pixels
 */
#ifndef TIF_B200_H
#define TIF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TIF_ABI_VERSION 5
#define TIF_MAX_FACES 128      /* 7-bit face id in a membership entry */
#define TIF_MAX_TANGENT_DIM 2048 /* 11-bit tangent row / column in a membership entry */

typedef enum TifStatus {
    TIF_OK = 0,
    TIF_ERR_INVALID_ARG = -1,
    TIF_ERR_CUDA = -2,          /* CUDA runtime error (including "no device") */
    TIF_ERR_ALLOC = -3,
    TIF_ERR_RANGE = -4,         /* plan: nearest tangent pixel fell outside the face, multiplicity > 7, ... */
    TIF_ERR_UNSUPPORTED = -5
} TifStatus;

/* face_blending_method of ico2erp_flow (projection_icosahedron.py:251,368-381) */
#define TIF_BLEND_STRAIGHTFORWARD 0
#define TIF_BLEND_NORMWARP 1
/* weights of cubemap2erp_flow (projection_cubemap.py:333-345, projection.py:64-117): 1 inside the unpadded
 * square / 0 outside, or 0.95 / ||tar(end point) - src||_2 on uint8-rounded samples */
#define TIF_BLEND_CUBEMAP_INSIDE 2
#define TIF_BLEND_CUBEMAP_WARP_ERROR 3

/* plan kinds (tif_plan_create flags): which reference stage the face table belongs to */
#define TIF_PLAN_ICOSAHEDRON 0     /* projection_icosahedron.py: sph2erp pixel centres, triangle faces */
#define TIF_PLAN_CUBEMAP 1         /* projection_cubemap.py: x = (theta+pi)/(2 pi) W resampling grid, square faces */

/* boundary rule of warp_backward (flow_warp.py:82-86): 3-D input -> scipy 'wrap' (period n-1),
 * 2-D input -> scipy 'constant' with cval 255 */
#define TIF_WARP_WRAP 0
#define TIF_WARP_CONSTANT255 1

#define TIF_DTYPE_F32 0
#define TIF_DTYPE_F64 1
#define TIF_DTYPE_U8 2           /* images of tif_blend_weight_points only */

/*
 * One tangent face.  Replaces the dict returned by get_icosahedron_parameters
 * (projection_icosahedron.py:19-160) plus the per-face scalars erp2ico_image (:199-208) and
 * ico2erp_flow (:285-305) derive from it.  Every double is computed on the host in float64 with the
 * reference's own operation order, so the device only adds, multiplies, divides and compares.
 */
typedef struct TifFace {
    double theta0, phi0;            /* tangent point */
    double sin_phi0, cos_phi0;      /* np.sin / np.cos of phi0 */
    double tri[4][2];               /* padded face polygon, gnomonic (x, y), reference vertex order (3 or 4 used) */
    double inner[4][2];             /* cubemap: the unpadded square of the "straightforward" weights (projection.py:87-93) */
    double xmin, xmax, ymin, ymax;  /* gnomonic bounding box of the padded triangle */
    double g2p_x, g2p_y;            /* (Wt-1)/(xmax-xmin), (Ht-1)/(ymax-ymin): gnomonic2pixel ratios
                                       (gnomonic_projection.py:141,145) */
    double p2g_x, p2g_y;            /* (Wt-1)/|xmax-xmin|, (Ht-1)/|ymax-ymin|: pixel2gnomonic ratios (:188,192) */
    double eps_stitch;              /* |xmax-xmin| / (2*Wt)   (projection_icosahedron.py:296) */
    double eps_image;               /* (xmax-xmin) / Wt       (:215, full_face_image=False) */
    double eps_inner;               /* cubemap: 1e-7 (projection.py:92) */
    int32_t col_start, col_stop;    /* integer ERP bounding box (:299-305); may leave the image */
    int32_t row_start, row_stop;
    int32_t tangent_h;              /* Ht of this face */
    int32_t pix_offset;             /* first pixel of this face in the tangent stack */
    int32_t n_vertices;             /* 3 (icosahedron) or 4 (cubemap) */
    int32_t n_inner;                /* 0, or 4 for the cubemap's inner square */
} TifFace;

/*
 * Host-side lookup tables handed to tif_plan_create (all HOST pointers, copied during the call).
 * They hold every transcendental the membership test needs, evaluated by the host's numpy exactly as
 * the reference evaluates them, which is what makes membership and pixel indices bit-identical.
 */
typedef struct TifPlanTables {
    const double* gnom_x;       /* [F][Wt]   np.linspace(xmin, xmax, Wt)      (projection_icosahedron.py:207) */
    const double* gnom_y;       /* [P / Wt]  np.linspace(ymax, ymin, Ht_f), faces concatenated (:208) */
    const double* row_sincos;   /* [H][2]    sin(phi_r), cos(phi_r), phi_r per spherical_coordinates.py:93 */
    const double* col_sincos;   /* [F][W][2] sin(theta_c - theta0_f), cos(theta_c - theta0_f) (gnomonic_projection.py:36) */
} TifPlanTables;

typedef struct TifPlan TifPlan;     /* opaque; owns device tables */

typedef struct TifPlanInfo {
    int32_t n_faces, erp_h, erp_w, tangent_w;
    int64_t tangent_pixels;         /* P */
    int32_t max_faces_per_pixel;    /* K: membership planes */
    int64_t accepted_unique;        /* sum over ERP pixels of accepting faces */
    int64_t referenced_tangent_pixels; /* unique tangent pixels the stitch reads (P_ref) */
    int64_t device_bytes;           /* plan footprint in HBM */
    int64_t stage1_patches;         /* 16 x 8 tangent patches of tif_erp2ico_u8 ... */
    int64_t stage1_staged_patches;  /* ... and how many of them are resampled through shared memory (0: ERP rows are
                                       not 16-byte multiples, every patch takes the direct gathers) */
    int64_t stage1_tma_patches;     /* ... and how many fit one of the two TMA box shapes below (cp.async.bulk.tensor) */
    int32_t stage1_tma_box_rows;    /* rows of both boxes */
    int32_t stage1_tma_box_bytes[2];/* row length of the narrow / the wide box */
    int32_t reserved;
} TifPlanInfo;

int tif_abi_version(void);
/* sizeof(TifFace), sizeof(TifPlanTables), sizeof(TifPlanInfo) as compiled into the library, so a
 * binding in another language can verify its struct layout before the first call. */
int tif_struct_sizes(int32_t* face_bytes, int32_t* tables_bytes, int32_t* info_bytes);
const char* tif_status_string(int status);
const char* tif_last_error(void);

/*
 * Geometry plan for one (face table, H, W, Wt) configuration: per tangent pixel the ERP sample
 * position (stage 1), per ERP pixel the accepting faces and nearest tangent pixels (stage 2).
 * Input-independent, built by fp64 kernels on the device, kept resident in HBM and reused by every
 * execute call.  Replaces the geometry half of erp2ico_image (projection_icosahedron.py:197-229)
 * and ico2erp_flow (:282-329), which the reference recomputes for every face of every call.
 */
int tif_plan_create(const TifFace* faces /*host*/, int n_faces, int erp_h, int erp_w, int tangent_w,
                    const TifPlanTables* tables /*host*/, int kind /*TIF_PLAN_*/, void* stream, TifPlan** plan_out);
int tif_plan_destroy(TifPlan* plan);
int tif_plan_info(const TifPlan* plan, TifPlanInfo* info /*host*/);

/* Parity / debug exports (device outputs).
 * stage 1: float ERP sample coordinate of each tangent pixel after sph2erp(modulo) (:225), and the
 *          inside-padded-triangle flag used by full_face_image=False (:214-218).
 * stage 2: per ERP pixel the number of accepting faces and, plane k, the packed entry
 *          face<<25 | multiplicity<<22 | iy<<11 | ix  (faces in ascending order).  In a cubemap plan the 3-bit
 *          field holds the inside-the-unpadded-square flag instead of the multiplicity. */
int tif_plan_export_stage1(const TifPlan* plan, double* erp_x /*[P]*/, double* erp_y /*[P]*/,
                           uint8_t* inside /*[P]*/, void* stream);
int tif_plan_export_stage2(const TifPlan* plan, uint8_t* count /*[H*W]*/, uint32_t* entries /*[K][H*W]*/,
                           void* stream);

/*
 * erp2ico_image (projection_icosahedron.py:163-248): ERP -> padded tangent images, bilinear with
 * scipy's mode='wrap' (period n-1) and uint8 round-half-up; RGB = samples, A = 255 (0 and RGB 255
 * outside the padded triangle when full_face_image == 0).
 * Any byte alignment of `erp` and any H are accepted; the fast (shared-memory staged) form needs a 16-byte aligned
 * `erp` and ERP rows of 6 H bytes that are 16-byte multiples (H % 8 == 0), see TifPlanInfo.stage1_staged_patches.
 */
int tif_erp2ico_u8(const TifPlan* plan, const uint8_t* erp /*[B,H,W,3]*/, int batch, int full_face_image,
                   uint8_t* tangent_rgba /*[B,P,4]*/, void* stream);

/* The same for float images of any channel count (`dtype` TIF_DTYPE_F32 / F64, [B,H,W,C]): scipy samples them in float64
 * and the reference casts the result into its integer tangent arrays (projection_icosahedron.py:231-241); `tangent` is
 * float64 [B,P,C], the cast is the caller's.  Pixels outside the padded triangle (full_face_image == 0) hold 255. */
int tif_erp2ico_float(const TifPlan* plan, const void* erp, int dtype, int batch, int channels, int full_face_image,
                      double* tangent, void* stream);

/*
 * Compact forms of a tangent stack for the trip back to the host (n_pixels = B * P, a multiple of 4):
 * TIF_PACK_RGB drops the alpha channel, which is a constant 255 for full_face_image (projection_icosahedron.py:244);
 * TIF_PACK_GRAY_CV is what the per-face estimator computes first, cv2.cvtColor(img[:, :, :3], cv2.COLOR_BGR2GRAY)
 * (flow_estimate.py:31-47), bit for bit: (3735 c0 + 19235 c1 + 9798 c2 + 2^14) >> 15.
 */
#define TIF_PACK_RGB 0
#define TIF_PACK_GRAY_CV 1
int tif_pack_tangent_u8(const uint8_t* tangent_rgba /*[n_pixels,4]*/, int64_t n_pixels, int format,
                        uint8_t* out /*[n_pixels,3] or [n_pixels]*/, void* stream);

/*
 * ico2erp_flow (projection_icosahedron.py:251-398): tangent flows -> ERP flow.  Nearest-pixel
 * gather, end-point lift through pixel2gnomonic / reverse_gnomonic_projection / sph2erp, seam fix
 * (wrap_around), face weights (blend), weighted mean over faces in index order, and the final
 * flow_postproc.erp_of_wraparound when wrap_around != 0.  src/tar are only read for
 * TIF_BLEND_NORMWARP (projection.py:15-61).  `scratch` (always required) must hold
 * tif_ico2erp_scratch_bytes(): one lifted record per tangent pixel and image, plus the normwarp face sums.
 */
int64_t tif_ico2erp_scratch_bytes(const TifPlan* plan, int batch, int blend);
int tif_ico2erp_flow_f32(const TifPlan* plan, const float* tangent_flow /*[B,P,2]*/,
                         const uint8_t* erp_src /*[B,H,W,3] or NULL*/, const uint8_t* erp_tar /*[B,H,W,3] or NULL*/,
                         int batch, int blend, int wrap_around, float* erp_flow /*[B,H,W,2]*/,
                         void* scratch, void* stream);

/* Profiling aid: the same call restricted to some of its kernels.  pass_mask bit 0: lift the tangent pixels and
 * (normwarp) accumulate the face-global warp-error sums; bit 1: blend using what is already in `scratch`;
 * bit 2: skip the lift kernel and reuse the records of an earlier call; bit 3 alone: the lift kernel only.
 * 3 = the call above. */
int tif_ico2erp_flow_f32_ex(const TifPlan* plan, const float* tangent_flow, const uint8_t* erp_src,
                            const uint8_t* erp_tar, int batch, int blend, int wrap_around, float* erp_flow,
                            void* scratch, int pass_mask, void* stream);

/*
 * flow_warp.warp_backward (flow_warp.py:54-88): out(y,x) = bilinear(img, (y+v, x+u)).
 * flow_dtype selects float32 / float64 flow ([B,H,W,2]); float64 reproduces the reference's
 * coordinates exactly.  uint8 images are rounded half-up like scipy; float images stay float.
 * uint8 RGB images with the 'wrap' boundary, W % 4 == 0 and 4-byte aligned buffers take a fast kernel (integer
 * coordinates + 23-bit fractions, aligned-word taps, 2^24 fixed-point blend, float64 only where the rounding is
 * undecided); its output is identical to the general float64 kernel's.
 */
int tif_warp_backward_u8(const uint8_t* image /*[B,H,W,C]*/, const void* flow, int flow_dtype,
                         int batch, int h, int w, int channels, int mode, uint8_t* out, void* stream);
int tif_warp_backward_f32(const float* image /*[B,H,W,C]*/, const void* flow, int flow_dtype,
                          int batch, int h, int w, int channels, int mode, float* out, void* stream);
/* float64 images stay float64 (scipy's own accumulation order), so the result is the reference's bit for bit */
int tif_warp_backward_f64(const double* image /*[B,H,W,C]*/, const void* flow, int flow_dtype,
                          int batch, int h, int w, int channels, int mode, double* out, void* stream);

/* flow_warp.flow_warp_meshgrid (flow_warp.py:15-51): end points with a single wrap; out [B,2,H,W]. */
int tif_flow_warp_meshgrid(const void* flow_u /*[B,H,W]*/, const void* flow_v /*[B,H,W]*/, int dtype,
                           int batch, int h, int w, void* out /*[B,2,H,W] same dtype*/, void* stream);

/*
 * Global rotation (row f2 of SURVEY.md section 8): the de-rotation steps either side of the tangent-image path.
 *
 * tif_rotation_motion_vector_f64   spherical_coordinates.rotation2erp_motion_vector (spherical_coordinates.py:185-218):
 *     ERP flow [H,W,2] float64 of a rigid rotation (`rotation` = 9 HOST doubles, row major).  rotate_erp_array
 *     (:172-182) is this with R^T followed by tif_warp_backward_* with the float64 flow.
 * tif_flow_cross_covariance        the 3x3 matrix S^T T of flow_warp.flow2rotation_3d (flow_warp.py:103-127,
 *     pointcloud_utils.py:29): unit vectors of the pixel centres and of the flow end points over rows
 *     [H/4, 3H/4).  `out` [B,9] float64 on the device; deterministic two-level reduction through `scratch`
 *     (tif_flow_cross_covariance_scratch_bytes).  The 3x3 SVD that follows is 9 numbers and stays with the caller.
 * tif_flow_rotate_endpoint         projection.flow_rotate_endpoint (projection.py:120-159): `rotation_field` is the
 *     [H,W,2] float64 output of tif_rotation_motion_vector_f64(R, wraparound=1).
 */
int tif_rotation_motion_vector_f64(const double* rotation /*host, 9*/, int h, int w, int wraparound,
                                   double* out /*[H,W,2]*/, void* stream);
int64_t tif_flow_cross_covariance_scratch_bytes(int batch, int h, int w);
/* flow2rotation_2d (flow_warp.py:136-187, use_weight=False): six float64 sums per image -- sum of u over the image;
 * over the columns [col_start, col_stop): sum of sign(v), sum and count of v > 0, sum and count of v < 0.  The
 * centre band depends on the first sum, so the caller runs it twice; fixed-order reduction, reproducible. */
int64_t tif_flow_rotation2d_stats_scratch_bytes(int batch, int h, int w);
int tif_flow_rotation2d_stats(const void* flow /*[B,H,W,2]*/, int dtype, int batch, int h, int w, int col_start,
                              int col_stop, double* out /*[B,6]*/, void* scratch, void* stream);
int tif_flow_cross_covariance(const void* flow /*[B,H,W,2]*/, int dtype, int batch, int h, int w, double* out /*[B,9]*/,
                              void* scratch, void* stream);
int tif_flow_rotate_endpoint(const void* flow /*[B,H,W,2]*/, int dtype, const double* rotation_field /*[H,W,2]*/,
                             int batch, int h, int w, int wraparound, void* out /*[B,H,W,2] same dtype*/, void* stream);

/*
 * Flow metrics (row f3): flow_evaluate.{available_pixel, EPE_mat, RMSE_mat, AAE_mat} (flow_evaluate.py:23-217).
 * Per-pixel error matrices of an evaluated flow against a ground truth (both [B,H,W,2], float32 or float64):
 * `epe` = end-point error (planar Euclidean, or great-circle distance of the two end points when spherical != 0),
 * `aae` (nullable) = planar angular error |atan2(cross, dot)|, `valid` = available_pixel of the ground truth
 * (and of `mask` [B,H,W] uint8 when given).  Where the ground truth is unusable both flows count as zero.
 * tif_reduce_f64 sums n doubles per image (power 1) or their squares (power 2) deterministically:
 * EPE = sum(epe)/sum(valid), RMSE = sqrt(sum(epe^2)/sum(valid)), AAE = sum(aae)/sum(valid).
 */
int tif_flow_error_mats(const void* gt, const void* eva, int dtype, const uint8_t* mask /*nullable*/, int batch, int h,
                        int w, int spherical, double* epe /*[B,H,W]*/, double* aae /*[B,H,W] nullable*/,
                        uint8_t* valid /*[B,H,W]*/, void* stream);
int64_t tif_reduce_f64_scratch_bytes(int batch);
int tif_reduce_f64(const double* x /*[B,n]*/, int batch, int64_t n, int power, double* out /*[B]*/, void* scratch,
                   void* stream);

/*
 * projection.get_blend_weight_ico / get_blend_weight_cubemap, weight_type "image_warp_error" (projection.py:15-117),
 * as a stand-alone call (tif_ico2erp_flow_f32 fuses the same weights): n source positions and n target positions in
 * ERP pixels, two [H,W,C] images (TIF_DTYPE_U8 / F32 / F64), scipy map_coordinates(order=1, mode='constant',
 * cval=255) per channel.  weight_kind 0: exp(-mean_c|tar - src| / mean over all points and channels);
 * weight_kind 1: 0.95 / ||tar - src||_2, 1 where that is 0.  Deterministic.
 */
int64_t tif_blend_weight_points_scratch_bytes(int64_t n);
int tif_blend_weight_points(const void* image_src, const void* image_tar, int image_dtype, int h, int w, int channels,
                            const double* x_src, const double* y_src, const double* x_tar, const double* y_tar,
                            int64_t n, int weight_kind, double* weights /*[n]*/, void* scratch, void* stream);

/*
 * Colour-wheel rendering of a flow field (row f4): flow_vis.flow_to_color / flow_uv_to_colors (flow_vis.py:77-201).
 * tif_flow_rad_max: max over the pixels of sqrt(u^2 + v^2) after the optional clip (`out`: one double on the device);
 * tif_flow_to_color_u8: clip, divide by `norm` (rad_max + 1e-5; 0 = already normalised), angle -> two rows of the
 * 55 x 3 wheel (`wheel`, device, from flow_vis.make_colorwheel), interpolate, floor(255 c); float64 arithmetic.
 */
int tif_flow_rad_max(const void* flow /*[n_pixels,2]*/, int dtype, int64_t n_pixels, int use_clip, double clip_lo,
                     double clip_hi, double* out, void* stream);
int tif_flow_to_color_u8(const void* flow /*[n_pixels,2]*/, int dtype, int64_t n_pixels, int use_clip, double clip_lo,
                         double clip_hi, double norm, const double* wheel /*[55,3]*/, int convert_to_bgr,
                         uint8_t* out /*[n_pixels,3]*/, void* stream);

/*
 * Collecting the stitched flows of several GPUs on one of them (the only exchange of the sharded path, SURVEY.md 8e;
 * the reference has no counterpart, it loops its pairs serially).  One process per GPU: the collecting process
 * allocates the [n_items, H, W, 2] buffer with tif_ipc_alloc and exports its CUDA IPC handle (TIF_IPC_HANDLE_BYTES
 * bytes, sent through any host channel); the others map it with tif_ipc_open and pass the device address of their own
 * slice as `erp_flow` to tif_ico2erp_flow_f32, whose blend epilogue then stores across NVLink / NVSwitch straight into
 * the collecting GPU's memory.  tif_ipc_copy_async moves a finished local result instead (cudaMemcpyAsync on `stream`).
 */
#define TIF_IPC_HANDLE_BYTES 64
int tif_ipc_alloc(void** dev_ptr, int64_t bytes);
int tif_ipc_free(void* dev_ptr);
int tif_ipc_export(void* dev_ptr, unsigned char* handle /*host, TIF_IPC_HANDLE_BYTES*/);
int tif_ipc_open(const unsigned char* handle /*host*/, void** dev_ptr);
int tif_ipc_close(void* dev_ptr);
int tif_ipc_copy_async(void* dst, const void* src, int64_t bytes, void* stream);

/* flow_postproc.erp_of_wraparound (flow_postproc.py:17-44), in place on [B,H,W,2]. */
int tif_erp_of_wraparound(void* erp_flow, int dtype, int batch, int h, int w, double u_threshold, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TIF_B200_H */
