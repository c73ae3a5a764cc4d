/*
 * hoc_b200.h -- C ABI of libhoc_b200.so: the sm_100a (B200) kernels behind the
 * differentiable-render + photometric-consistency path of hassony2/handobjectconsist.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes, no torch / C++ types; every pointer is a DEVICE pointer to a
 *     contiguous buffer owned by the caller unless the parameter name ends in `_host`;
 *   - fp32 data, int32 index maps, int64 not used;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream); launches
 *     are asynchronous and allocate nothing -- scratch memory is caller-provided, its size
 *     comes from the matching `*_workspace_bytes` query;
 *   - return value: HOC_OK (0) or a negative HOC_ERR_* code; `hoc_last_error()` returns a
 *     thread-local human-readable message.  No exception crosses the boundary.
 *
 * Each entry point names the reference interface it replaces.  The reference's rasterizer
 * kernels live in the third-party extension `neural_renderer.cuda.rasterize` and are bound at
 * /root/reference/meshreg/neurender/rasterize.py:202,232,269,290,306; the warp / loss path is
 * /root/reference/meshreg/warping/imgflowarp.py and meshreg/optim/{pyramidloss,lossutils}.py;
 * MANO skinning is `manopth.manolayer.ManoLayer.forward`, bound at
 * /root/reference/meshreg/models/manobranch.py:70-85,101-145.
 */
#ifndef HOC_B200_H
#define HOC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HOC_ABI_VERSION 1

#define HOC_OK 0
#define HOC_ERR_INVALID_ARG (-1)
#define HOC_ERR_CUDA (-2)
#define HOC_ERR_WORKSPACE (-3)

/* Output layouts of the rasterizer.
 * RAW   = what RasterizeFunction.forward returns (rasterize.py:118-125): rgb [B,S,S,3],
 *         alpha/depth [B,S,S], raster row order (row 0 is the image BOTTOM).
 * IMAGE = what rasterize_rgbad returns after its permute + row flip (rasterize.py:417-428):
 *         rgb [B,3,S,S], alpha/depth [B,S,S] with row 0 at the image TOP.  Writing this layout
 *         directly removes the three gather kernels of the reference.
 * face_index_map / weight_map / face_inv_map are never flipped (rasterize.py:443-445). */
#define HOC_LAYOUT_RAW 0
#define HOC_LAYOUT_IMAGE 1
/* OR-ed into `layout` of hoc_raster_forward: the workspace already holds 0xff bytes (hoc_mesh_gather_clear did it),
 * so the forward skips its own fill -- one graph node less on the critical path of the fused flow path */
#define HOC_LAYOUT_KEYS_CLEARED 0x100
/* OR-ed into `layout` of hoc_raster_forward: `textures` is [B,F,3,3] -- three vertex values c0, c1, c2 per face -- and
 * stands for the cube T[i,j,k] = i c0 + j c1 + k c2 (texture_size 2), whose texels the forward evaluates on the fly
 * with hoc_mesh_gather's expression: same bits as sampling the materialised cube, 36 instead of 96 bytes per face */
#define HOC_LAYOUT_TEX_VERTEX 0x200
/* OR-ed into `layout` of hoc_raster_forward: `depth` and `weight_map` are only kept for hoc_raster_backward, which
 * reads them at covered pixels -- write them there only (background pixels stay uninitialised): 16 of the 36 bytes per
 * pixel the forward writes are not written for the 90+ % of the pixels no face covers */
#define HOC_LAYOUT_SPARSE_SAVED 0x400

int hoc_abi_version(void);
const char *hoc_last_error(void);

/* Kernel ids for launch accounting / device timing (bench.py's `gpu_launches` and `roofline`). */
#define HOC_K_RASTER_ZBUF 0
#define HOC_K_RASTER_RESOLVE 1
#define HOC_K_GRAD_EXTENT 2 /* retired: merged into HOC_K_RASTER_BWD_PIXEL */
#define HOC_K_RASTER_BACKWARD 3 /* hoc_raster_bwd_depth_kernel (per-face epilogue of backward_depth_map) */
#define HOC_K_WARP_PHOTO_FWD 4
#define HOC_K_WARP_PHOTO_BWD 5
#define HOC_K_WARP 6
#define HOC_K_WARP_BWD 7
#define HOC_K_OCCLUSION 8
#define HOC_K_MESH_GATHER 9
#define HOC_K_MESH_SCATTER 10
#define HOC_K_FLOW_FINALIZE 11
#define HOC_K_FLOW_FINALIZE_BWD 12
#define HOC_K_RASTER_BWD_PIXEL 13 /* hoc_raster_bwd_scan_kernel (streaming pass) */
#define HOC_K_RASTER_BWD_LINE 14
#define HOC_K_FLOW_VERTICES 15
#define HOC_K_FLOW_VERTICES_BWD 16
#define HOC_K_MANO_FWD 17
#define HOC_K_MANO_BWD 18
#define HOC_K_RASTER_BWD_PIXEL_K4 19 /* (rounds 1-2: the cover pass with the pseudo-gradient's per-pixel work; unused) */
#define HOC_K_RASTER_BACKWARD_COVER 20 /* hoc_raster_bwd_cover_kernel<.., false>: texture / depth gradient only */
#define HOC_K_CAT_MESHES 21
#define HOC_K_PAIR_LOSS 22
#define HOC_K_UNPACK_U8 23
#define HOC_K_HAND_HEAD_FWD 24
#define HOC_K_HAND_HEAD_BWD 25
#define HOC_K_RECOVER_POINTS_FWD 26
#define HOC_K_RECOVER_POINTS_BWD 27
#define HOC_K_PAIR_FRONT 28
#define HOC_K_PAIR_BACK 29
#define HOC_K_AUGMENT_STATS 30
#define HOC_K_AUGMENT_FRAMES 31
#define HOC_K_RASTER_BWD_GROUP 32 /* not a kernel: the whole hoc_raster_backward(_ex) call (scan + cover + line [+ depth]) timed
                                     with ONE event pair -- per-kernel event nodes add ~4 us each inside a captured graph */
#define HOC_KERNEL_COUNT 33

/* Tuning knobs (defaults are the measured optimum on B200; meant for benchmarking sweeps).
 *   HOC_TUNE_LINE_THREADS  threads per CTA of the rasterizer backward's line pass (multiple of 32, <= 256)
 *   HOC_TUNE_LINE_SEGMENT  pixels per work item of an outward scan (8, 16 or 32; 0 = by raster size, the default)
 *   HOC_TUNE_LINE_LINES / HOC_TUNE_LINE_FOLD  image lines per CTA of the line pass and the way they are dealt (below)
 *   HOC_TUNE_DETERMINISTIC 0 (default) / 1: reproducible mode.  The gradient sums that production accumulates with
 *                          float atomics (like the reference's backward_textures / backward_depth_map /
 *                          index_put(accumulate)) are accumulated in 128-bit fixed point with integer atomics
 *                          instead (csrc/hoc_det.cuh): the result no longer depends on the order of the additions,
 *                          so two runs -- or a captured graph and the eager path -- give the same bits.  Needs the
 *                          larger workspaces reported by hoc_raster_backward_workspace_bytes_ex /
 *                          hoc_mesh_scatter_workspace_bytes.  Slower; for tests and debugging. */
#define HOC_TUNE_LINE_THREADS 1
#define HOC_TUNE_LINE_SEGMENT 2
#define HOC_TUNE_DETERMINISTIC 3
#define HOC_TUNE_PDL 5       /* 0 (default) / 1: programmatic dependent launch of the frame-pair step's kernels (launch,
                              * CTA scheduling and prologue of kernel N + 1 overlap the tail of kernel N; results are
                              * identical either way).  Measured: no gain inside the captured graph */
#define HOC_TUNE_TEX_IN_LINE 9 /* 1 (default) / 0: the line pass's row CTAs also run backward_textures (vertex-value
                                * textures, saved weights: the frame-pair path); 0 = a cover pass does it */
#define HOC_TUNE_COVER_CTAS 6 /* CTAs per sample of the rasterizer backward's cover pass (grid-stride over the listed pixels) */
#define HOC_TUNE_LINE_LINES 7 /* lines (image columns / rows) per CTA of the line pass, 1 .. 8; 0 (default) = by raster
                               * size: 1 up to 320, 3 above.  Line k of CTA y has centre-out rank k * G + y, G = CTAs per
                               * (sample, axis) */
#define HOC_TUNE_LINE_FOLD 8  /* 1 (default) / 0: odd k take rank k * G + (G - 1 - y) instead -- a heavy central line
                               * shares its CTA with a light outer one */
int hoc_set_tuning(int key, int value);

/* Number of launches of one kernel (or of all kernels, kernel_id = -1) since the library was loaded. */
unsigned long long hoc_launch_count(int kernel_id);
/* Arm / read the device timer: between begin and end every launch of a kernel whose bit is set in
 * `kernel_mask` (1 << HOC_K_*) is bracketed by CUDA events on its stream; end synchronises them and writes
 * up to `capacity` durations (ms) and kernel ids to host memory, returning how many were recorded.  Not
 * thread-safe; meant for benchmarking. */
int hoc_timer_begin(unsigned long long kernel_mask);
int hoc_timer_end(float *ms_host, int *kernel_ids_host, int capacity);
/* Timing kernels INSIDE a CUDA graph: arm the timer, capture the graph (the brackets become external event
 * nodes that every replay re-records), call hoc_timer_pause() (stops bracketing, keeps the event pairs; returns
 * how many), replay, then hoc_timer_peek() reads the durations of the latest replay without resetting. */
int hoc_timer_pause(void);
int hoc_timer_peek(float *ms_host, int *kernel_ids_host, int capacity);

/* ---- rasterizer forward -------------------------------------------------------------------
 * Replaces forward_face_index_map + forward_texture_sampling (rasterize.py:202-215,232-243)
 * plus the wrapper's fill_/background/alpha/clone/flip ops (rasterize.py:58-125,246-260,417-428).
 *   faces      [B,F,3,3]  x,y in NDC (y up), z metric
 *   textures   [B,F,ts,ts,ts,3] or NULL (then rgb must be NULL)
 *   background_host  3 floats (host), used when background_dev is NULL
 *   background_dev   [B,3] per-sample background or NULL
 *   rgb / alpha / depth            outputs in `layout`, any of them may be NULL (not produced)
 *   face_index_map [B,S,S] int32   (-1 = background)                        never NULL
 *   weight_map     [B,S,S,3]       may be NULL
 *   face_inv_map   [B,S,S,3,3]     may be NULL (the reference fills it only when return_depth)
 *   workspace      hoc_raster_forward_workspace_bytes(B,F,S) bytes
 */
size_t hoc_raster_forward_workspace_bytes(int B, int F, int S);
int hoc_raster_forward(const float *faces, const float *textures, int B, int F, int S, int ts, float near_,
                       float far_, float eps, const float *background_host, const float *background_dev,
                       int layout, float *rgb, float *alpha, float *depth, int32_t *face_index_map,
                       float *weight_map, float *face_inv_map, void *workspace, size_t workspace_bytes,
                       void *stream);
/* The same with a raster row window (device int [B] or NULL, see hoc_pair_front): rows yi < row_lo[b] of sample b are
 * not drawn and -- by the 4-pixel image-layout resolve pass the frame-pair path uses -- not written (undefined). */
int hoc_raster_forward_ex(const float *faces, const float *textures, int B, int F, int S, int ts, float near_,
                          float far_, float eps, const float *background_host, const float *background_dev,
                          int layout, const int *row_lo, float *rgb, float *alpha, float *depth,
                          int32_t *face_index_map, float *weight_map, float *face_inv_map, void *workspace,
                          size_t workspace_bytes, void *stream);

/* ---- rasterizer backward ------------------------------------------------------------------
 * Replaces backward_pixel_map + backward_textures + backward_depth_map
 * (rasterize.py:269-281,290-297,306-315) and the zero-fills around them (rasterize.py:151-181).
 * Two launches -- a streaming scan pass and a line pass that runs the pseudo-gradient of one image column / row per CTA
 * (and the texture gradient, for vertex-value textures with saved weights) --, plus a cover pass for cube textures /
 * depth gradients and a per-face epilogue for the depth gradient (csrc/raster_bwd.cu).  Texture / depth gradients and
 * the pseudo-gradient are accumulated with float atomics, like the reference's backward_textures /
 * backward_depth_map: results are reproducible to rounding, not bit for bit (HOC_TUNE_DETERMINISTIC: bit for bit).
 *   faces, textures, face_index_map   forward inputs / output
 *   rgb              forward output in `layout` (NULL when the forward had no rgb)
 *   weight_map, depth  forward outputs ([B,S,S,3] raster order / [B,S,S] in `layout`) or NULL: when both are
 *                    given the pixel pass reads them instead of recomputing weights and depth from the faces
 *   grad_rgb / grad_alpha / grad_depth  incoming gradients in `layout`; NULL = all zeros / that
 *                    output was not requested in the forward
 *   use_alpha        1 when the forward produced alpha (return_alpha), else 0
 *   grad_faces       [B,F,3,3] out (fully overwritten) or NULL to skip the geometry gradient
 *   grad_textures    [B,F,ts,ts,ts,3] out (fully overwritten) or NULL to skip it
 *   workspace        hoc_raster_backward_workspace_bytes(B,F,S) bytes: line spans, counters, per-face depth sums and
 *                    the list of pixels with a texture / depth gradient (8 S^2 B, of which only the used part is
 *                    ever touched)
 */
/* tex_grad_mode: CUBE = grad_textures is [B,F,ts,ts,ts,3] (the reference's backward_textures);
 * VERTEX = the cubes were built by hoc_mesh_gather from three vertex values per face (ts == 2): grad_textures
 * is [B,F,3,3] = d loss / d (vertex k of the face, channel c) -- nine sums per face instead of twenty-four. */
#define HOC_TEX_GRAD_CUBE 0
#define HOC_TEX_GRAD_VERTEX 1
size_t hoc_raster_backward_workspace_bytes(int B, int F, int S);
/* the same, plus (reproducible mode only) the fixed-point accumulators for this texture size / gradient mode */
size_t hoc_raster_backward_workspace_bytes_ex(int B, int F, int S, int ts, int tex_grad_mode);
int hoc_raster_backward(const float *faces, const float *textures, const int32_t *face_index_map,
                        const float *rgb, const float *weight_map, const float *depth, const float *grad_rgb,
                        const float *grad_alpha,
                        const float *grad_depth, int B, int F, int S, int ts, float near_, float far_,
                        float eps, int layout, int use_alpha, int tex_grad_mode, float *grad_faces,
                        float *grad_textures, void *workspace, size_t workspace_bytes, void *stream);

/* The same for a batch of which only the first `geom_samples` samples need the pseudo-gradient
 * (backward_pixel_map); the rows of grad_faces of the others hold the depth gradient alone.  Used by the frame-pair
 * path, which stacks the two renders of a pair ([2B]) and differentiates the geometry of the first one only.
 * flags: HOC_BWD_WORKSPACE_ZEROED = the first hoc_raster_backward_zero_bytes(B,F,S) bytes of the workspace are
 * already zero (an earlier kernel of the caller's sequence filled them: one memset node less in a captured step).
 * extra_zero: a float buffer the streaming pass also zero-fills (the outputs of the hoc_mesh_scatter that follows).
 * row_lo: device int [B] or NULL -- the raster row window of hoc_pair_front (rows below it are not read). */
#define HOC_BWD_WORKSPACE_ZEROED 1
size_t hoc_raster_backward_zero_bytes(int B, int F, int S);
int hoc_raster_backward_ex(const float *faces, const float *textures, const int32_t *face_index_map,
                           const float *rgb, const float *weight_map, const float *depth, const float *grad_rgb,
                           const float *grad_alpha, const float *grad_depth, int B, int F, int S, int ts,
                           float near_, float far_, float eps, int layout, int use_alpha, int tex_grad_mode,
                           int geom_samples, int flags, void *extra_zero, size_t extra_zero_bytes, const int *row_lo,
                           float *grad_faces, float *grad_textures, void *workspace, size_t workspace_bytes,
                           void *stream);

/* ---- flow-guided warp + masked photometric L1 ---------------------------------------------
 * ONE direction of pair_consist (imgflowarp.py:80-107) in one launch: warp(src, flow),
 * warp(jitter, flow), the valid-mask algebra and criterion.compute / batch_masked_mean_loss
 * (pyramidloss.py:56-58, lossutils.py:1-8).  For the reference's direction 1 the caller passes
 * src = image_ref, target = image, flow = recons_flow[1], jitter = jitter_mask; for direction 2
 * src = image, target = image_ref, flow = recons_flow[0], jitter = jitter_mask_ref.
 *   src, target [B,C,H,W];  flow [B,H,W,2] pixel offsets;  jitter [B,Cj,H,W] (Cj == 1 or C) or NULL
 * outputs (any may be NULL except sums):
 *   warped     [B,C,H,W]  warp(src, flow), already multiplied by its in-bounds mask
 *   warp_mask  [B,C,H,W]  in-bounds mask * (warp(jitter, flow) == 1)
 *   valid_mask [B,H,W] uint8 (0/1)   warp_mask[:,0] & (flow[...,0] != 0) & (jitter[:,0] == 1)
 *   flow_mask  [B,H,W,2] uint8 (0/1) ~(flow == 0)
 *   diff       [B,C,H,W]  |warped - target|
 *   sums       [B,2] DOUBLE, zero-filled by the call; consumed by hoc_pair_loss / hoc_warp_photo_backward:
 *              (sum of diff over valid elements in units of 2^-28, number of valid elements) -- both integer-valued,
 *              so that the double atomics that accumulate them commute and the loss is reproducible bit for bit
 *   loss       [B] = sum / max(count, 1)  (batch_masked_mean_loss; second tiny launch; may be NULL)
 */
int hoc_warp_photo_forward(const float *src, const float *target, const float *flow, const float *jitter, int B,
                           int C, int Cj, int H, int W, float thresh, float *warped, float *warp_mask,
                           uint8_t *valid_mask, uint8_t *flow_mask, float *diff, double *sums, float *loss,
                           void *stream);

/* Same, but ACCUMULATES into `sums` (the caller zero-filled it, e.g. ahead of time on another stream). */
int hoc_warp_photo_forward_acc(const float *src, const float *target, const float *flow, const float *jitter, int B,
                               int C, int Cj, int H, int W, float thresh, float *warped, float *warp_mask,
                               uint8_t *valid_mask, uint8_t *flow_mask, float *diff, double *sums, float *loss,
                               void *stream);

/* Input side of the path: frames and jitter masks can cross PCIe as uint8 (a quarter of the fp32 bytes)
 * and be widened on the device: dst[i] = src[i] / div - sub, the two IEEE operations of torchvision's `to_tensor`
 * followed by `normalize(mean, 1)` (/root/reference/meshreg/datasets/handobjset.py:368-372: div = 255, sub = 0.5);
 * div = 255, sub = 0 for the jitter masks, `to_tensor` of a warped white image (handobjset.py:375-379). */
int hoc_unpack_u8(const uint8_t *src, float *dst, long long n, float div, float sub, void *stream);

/* BOTH directions of pair_consist (imgflowarp.py:80-107) in one launch, four pixels per thread (W % 4 == 0, C = 3,
 * 3-channel jitter masks or none; 16-byte aligned tensors).  Direction 0 warps image_ref with flow21 against image
 * (jitter mask of the second frame), direction 1 warps image with flow12 against image_ref; per-direction outputs are
 * passed as arrays of two pointers [direction 0, direction 1].
 *   visuals = 1  everything pair_consist returns: warped / warp_mask / diff [B,3,H,W] (any may be NULL)
 *   visuals = 0  training: only what the loss and its gradient need -- valid_mask, flow_mask, sums.  A pixel whose
 *                rendered flow is zero is invalid whatever it samples (imgflowarp.py:93-101), so it costs its flow
 *                and mask bytes only.
 *   sums [2,B,2] DOUBLE (direction-major), zero-filled by the call; feed hoc_pair_loss(sums, sums + 2 B, ...). */
int hoc_warp_photo_forward_pair(const float *image_ref, const float *image, const float *flow12, const float *flow21,
                                const float *jitter_ref, const float *jitter, int B, int H, int W, float thresh,
                                int visuals, float *const *warped, float *const *warp_mask, float *const *diff,
                                uint8_t *const *valid_mask, uint8_t *const *flow_mask, double *sums, void *stream);
/* Backward of both directions fused with hoc_flow_finalize_backward: grad_rgb1 / grad_rgb2 [B,3,S,S] (image layout,
 * fully overwritten: d loss / d flow x mult inside the H x W crop, zero elsewhere and in the third channel) are the
 * incoming gradients of the two renders; grad_flow12 / grad_flow21 [B,H,W,2] optionally receive d loss / d flow
 * itself.  Any output may be NULL.  use_backward = 0: direction 1 carries no loss (zeros).
 * grad_loss [B] and / or grad_mean [1]: d L / d loss[b] = grad_loss[b] + grad_mean / B (the adjoint of the batch mean
 * hoc_pair_loss_mean emits).  zero: an optional buffer the kernel also zero-fills (the counters of the
 * hoc_raster_backward_ex that follows, see HOC_BWD_WORKSPACE_ZEROED).  row_lo1 / row_lo2: raster row windows of the
 * two renders (device int [B] each, or NULL): rows of grad_rgb below them are not written. */
int hoc_warp_photo_backward_pair(const float *image_ref, const float *image, const float *flow12, const float *flow21,
                                 const uint8_t *const *valid_mask, const double *sums, const float *mult1,
                                 const float *mult2, const float *grad_loss, const float *grad_mean, int B, int S,
                                 int H, int W, int use_backward, float *grad_rgb1, float *grad_rgb2,
                                 float *grad_flow12, float *grad_flow21, void *zero, size_t zero_bytes,
                                 const int *row_lo1, const int *row_lo2, void *stream);
/* hoc_flow_finalize and the training half of hoc_warp_photo_forward_pair (visuals = 0) in ONE pass: the pixel that
 * has just produced its flow vector is the pixel whose warp that flow drives.  Same arguments as the two calls
 * (valid_mask / flow_mask / sums indexed by pair_consist's direction); `sums` must be zero on entry
 * (hoc_pair_front's zero region). */
int hoc_flow_finalize_warp(const float *rgb1, const float *alpha1, const int32_t *idx1, const float *rgb2,
                           const float *alpha2, const int32_t *idx2, const float *image_ref, const float *image,
                           const float *jitter_ref, const float *jitter, int B, int S, int H, int W,
                           const int *ignore_faces, int n_ignore, float distance_thresh, float thresh, float *flow12,
                           float *flow21, float *mult1, float *mult2, uint8_t *const *valid_mask,
                           uint8_t *const *flow_mask, double *sums, void *stream);
/* The same; sparse_outputs = 1: the step wants the loss and its gradient only -- flow12 / flow21 / mult1 / mult2 are
 * written where valid_mask is set (all that hoc_pair_backward_raster reads) and left undefined elsewhere: 24 B/px of
 * zero stores less (pass flow_mask = NULL as well). */
int hoc_flow_finalize_warp_ex(const float *rgb1, const float *alpha1, const int32_t *idx1, const float *rgb2,
                              const float *alpha2, const int32_t *idx2, const float *image_ref, const float *image,
                              const float *jitter_ref, const float *jitter, int B, int S, int H, int W,
                              const int *ignore_faces, int n_ignore, float distance_thresh, float thresh, float *flow12,
                              float *flow21, float *mult1, float *mult2, uint8_t *const *valid_mask,
                              uint8_t *const *flow_mask, double *sums, int sparse_outputs, void *stream);
/* hoc_pair_loss plus the mean over the batch (warpbranch.py:88 for one pair), one launch.  zero: an optional buffer the
 * launch also zero-fills (the counters of the step's rasterizer backward, HOC_BWD_WORKSPACE_ZEROED). */
int hoc_pair_loss_mean(const double *sums_fwd, const double *sums_bwd, int B, float *loss, float *mean, void *zero,
                       size_t zero_bytes, void *stream);
/* The backward of the frame-pair step from d loss to grad_faces / grad_textures in ONE call: the backward of pair_consist
 * fused into the rasterizer backward's scan pass (the incoming gradient of the renders' rgb maps is computed from the
 * valid masks instead of being written by one pass and read back by the next), then the line pass (pseudo-gradient of
 * the first geom_samples rows; texture gradient of all rows).  Two launches.
 * Arguments: those of hoc_warp_photo_backward_pair (pairs = B of the pair kernels) and of hoc_raster_backward_ex for the
 * stacked rows [row_offset, row_offset + n) of (render 1 of every pair, render 2 of every pair); image layout, vertex
 * texture gradients; grad_rgb [n,3,S,S] is scratch of the call (its third plane -- the rendered flow has two channels --
 * is not touched). */
size_t hoc_pair_backward_zero_bytes(int n, int F, int S); /* leading workspace bytes HOC_BWD_WORKSPACE_ZEROED vouches for */
int hoc_pair_backward_raster(const float *image_ref, const float *image, const float *flow12, const float *flow21,
                             const uint8_t *const *valid_mask, const double *sums, const float *mult1,
                             const float *mult2, const float *grad_loss, const float *grad_mean, int pairs, int H, int W,
                             int use_backward, int row_offset, const float *faces, const int32_t *face_index_map,
                             const float *rgb, const float *weight_map, const float *depth, float *grad_rgb, int n, int F,
                             int S, float near_, float far_, float eps, int geom_samples, int flags, void *extra_zero,
                             size_t extra_zero_bytes, const int *row_lo, float *grad_faces, float *grad_textures,
                             void *workspace, size_t workspace_bytes, void *stream);

/* pair_consist's per-sample loss from the sums of its two directions (imgflowarp.py:108-114):
 * loss[b] = masked_mean(bwd) + masked_mean(fwd) (that order, float) when sums_bwd is given, else masked_mean(fwd). */
int hoc_pair_loss(const double *sums_fwd, const double *sums_bwd, int B, float *loss, void *stream);

/* Gradient of loss[b] w.r.t. flow, the only differentiable input (imgflowarp.py:52-53: the
 * thresholded masks carry no gradient).  grad_loss [B]; valid_mask / sums from the forward;
 * grad_flow [B,H,W,2] out (fully overwritten). */
int hoc_warp_photo_backward(const float *src, const float *target, const float *flow, const uint8_t *valid_mask,
                            const double *sums, const float *grad_loss, int B, int C, int H, int W, float thresh,
                            float *grad_flow, void *stream);

/* Plain warp (imgflowarp.py:31-55).  flow_nchw [B,2,H,W]; mode 0 bilinear, 1 nearest.
 *   out [B,C,H,W] = grid_sample(x) * mask;  mask [B,C,H,W] or NULL. */
int hoc_warp(const float *x, const float *flow_nchw, int B, int C, int H, int W, float thresh, int mode,
             float *out, float *mask, void *stream);

/* Gradient of hoc_warp (bilinear) w.r.t. the flow: grad_out [B,C,H,W] is the gradient of the masked
 * output; grad_flow_nchw [B,2,H,W] out (fully overwritten).  The sampled image gets no gradient. */
int hoc_warp_backward(const float *x, const float *flow_nchw, const float *grad_out, int B, int C, int H, int W,
                      float thresh, float *grad_flow_nchw, void *stream);

/* Forward-backward occlusion check on rendered flows (get_occlusion_mask +
 * occlusion_mask_from_warped_grid, imgflowarp.py:118-172), both directions in one launch.
 *   mask1, mask2 [B,H,W];  flow12, flow21 [B,Cf,H,W] (first two channels used);
 *   occl1, occl2 [B,H,W] out. */
int hoc_occlusion_mask(const float *mask1, const float *mask2, const float *flow12, const float *flow21, int B,
                       int Cf, int H, int W, float distance_thresh, float *occl1, float *occl2, void *stream);

/* ---- mesh-flow glue (opticalflow.py:98-154, renderer.py:250-252,282) as kernels ---------------
 * hoc_mesh_gather: what batch_vertex_textures + fill_back + vertices_to_faces build with ~15 tensor ops.
 *   verts [B,V,3] (NDC x,y + metric z), attrs [B,V,3] per-vertex values (NULL: no textures),
 *   faces_idx [B,F,3] int64 -> faces_out [B,F',3,3], textures_out [B,F',2,2,2,3] (F' = 2F when fill_back:
 *   faces F..2F-1 are the reversed windings, their cubes the permute(0,1,4,3,2,5) of the originals). */
int hoc_mesh_gather(const float *verts, const float *attrs, const long long *faces_idx, int B, int V, int F,
                    int fill_back, float *faces_out, float *textures_out, void *stream);
/* hoc_mesh_gather that also fills `clear` (clear_bytes, a multiple of 16, 16-byte aligned; may be NULL) with 0xff
 * bytes: the z-buffer workspace of the hoc_raster_forward call that follows (pass HOC_LAYOUT_KEYS_CLEARED there).
 * tex_mode HOC_TEX_GRAD_VERTEX: textures_out is [B,F',3,3], the three vertex values of every face (for
 * HOC_LAYOUT_TEX_VERTEX) instead of the [B,F',2,2,2,3] cubes. */
int hoc_mesh_gather_clear(const float *verts, const float *attrs, const long long *faces_idx, int B, int V, int F,
                          int fill_back, int tex_mode, float *faces_out, float *textures_out, void *clear,
                          size_t clear_bytes, void *stream);
/* batch_cat_meshes (libyana.renderutils.catmesh, called at /root/reference/meshreg/models/warpbranch.py:50-52) for
 * the hand + object pair of one or two frames in ONE launch: verts_x [B,Vh+Vo,3] = cat(hand_x, obj_x),
 * faces [B,Fh+Fo,3] = cat(hand_faces, obj_faces + Vh).  hand_faces is [Fh,3] (shared) or [B,Fh,3]
 * (`hand_faces_batched`); verts_b / faces may be NULL to skip them. */
int hoc_cat_meshes(const float *hand_a, const float *obj_a, const float *hand_b, const float *obj_b,
                   const long long *hand_faces, int hand_faces_batched, const long long *obj_faces, int B, int Vh,
                   int Vo, int Fh, int Fo, float *verts_a, float *verts_b, long long *faces, void *stream);
/* Adjoint: grad_faces [B,F',3,3] / grad_textures (either may be NULL together with its output)
 * -> grad_verts [B,V,3], grad_attrs [B,V,3] (zero-filled by the call, accumulated with atomics).
 * tex_grad_mode as in hoc_raster_backward: CUBE = grad_textures [B,F',2,2,2,3], VERTEX = [B,F',3,3]. */
int hoc_mesh_scatter(const float *grad_faces, const float *grad_textures, const long long *faces_idx, int B, int V,
                     int F, int fill_back, int tex_grad_mode, float *grad_verts, float *grad_attrs, void *stream);
/* The same with a workspace: 0 bytes in production, the fixed-point accumulators of both outputs in the
 * reproducible mode (HOC_TUNE_DETERMINISTIC), where hoc_mesh_scatter itself fails with HOC_ERR_WORKSPACE.
 * outputs_zeroed = 1: the caller (an earlier kernel of its sequence) already filled both outputs with zeros.
 * Precondition of both (and of hoc_mesh_gather): 0 <= faces_idx < V; indices outside that range are skipped
 * (gather: read as vertex 0) instead of touching memory out of bounds. */
size_t hoc_mesh_scatter_workspace_bytes(int B, int V);
int hoc_mesh_scatter_ws(const float *grad_faces, const float *grad_textures, const long long *faces_idx, int B, int V,
                        int F, int fill_back, int tex_grad_mode, float *grad_verts, float *grad_attrs,
                        int outputs_zeroed, void *workspace, size_t workspace_bytes, void *stream);
/* Frame-pair front end in ONE launch (replaces hoc_cat_meshes + hoc_flow_vertices + 2 x hoc_mesh_gather_clear):
 * hand / object vertices of both frames [B,Vh,3] / [B,Vo,3] (camera space), hand_faces [Fh,3] (or [B,Fh,3]),
 * obj_faces [B,Fo,3] (object-local indices) and the cameras -> the rasterizer inputs of BOTH renders stacked along the
 * batch: faces_out / textures_out [2B,F',3,3] (F' = 2 (Fh + Fo) with fill_back; textures are the three vertex values
 * [dx, dy, 1] of HOC_LAYOUT_TEX_VERTEX), rows 0..B-1 = render of mesh 1 with flow 1->2, rows B..2B-1 = render of
 * mesh 2 with flow 2->1.  face_table [2B,Fh+Fo,3] (optional) receives the concatenated table the adjoint walks;
 * `clear` as in hoc_mesh_gather_clear (the z-buffer keys of the [2B] forward that follows); `zero` an optional buffer
 * filled with zeros (the loss sums of hoc_flow_finalize_warp).
 * row_lo (device int [2B], optional): the RASTER ROW WINDOW of every pair (warpreg.py:29,40-45 renders
 * the square that contains the frame, opticalflow.py:152-154 crops afterwards).  Raster rows yi < row_lo[b] (rows count
 * from the bottom; the crop keeps the top crop_h rows of the S x S raster) are skipped by hoc_raster_forward_ex,
 * hoc_warp_photo_backward_pair and hoc_raster_backward_ex when they are handed the same array.  The window is exact,
 * computed on the device from the pair itself: it extends below the crop by what the forward-backward occlusion check
 * can reach (two hops of the largest vertex displacement) and, with geom_window = 1 (the pseudo-gradient is wanted), down
 * to the lowest vertex of both meshes.
 * Replaces warpbranch.py:50-52, opticalflow.py:98-103,121-123, renderer.py:250-252,282. */
int hoc_pair_front(const float *hand1, const float *obj1, const float *hand2, const float *obj2,
                   const long long *hand_faces, int hand_faces_batched, const long long *obj_faces, const float *K1,
                   int K1_batched, const float *K2, int K2_batched, const float *R, int R_batched, const float *t,
                   int t_batched, const float *dist_coeffs, int dist_batched, float orig_size, int B, int Vh, int Vo,
                   int Fh, int Fo, int fill_back, float *faces_out, float *textures_out, long long *face_table,
                   void *clear, size_t clear_bytes, void *zero, size_t zero_bytes, int *row_lo, int S, int crop_h,
                   int geom_window, void *stream);
/* Its per-vertex adjoint: grad_ndc / grad_attrs [2B,Vh+Vo,3] (hoc_mesh_scatter's outputs for the stacked batch; the
 * has_* flags say which halves carry a gradient) -> grad_verts1 / grad_verts2 [B,Vh+Vo,3] (either may be NULL). */
int hoc_pair_back(const float *hand1, const float *obj1, const float *hand2, const float *obj2, const float *K1,
                  int K1_batched, const float *K2, int K2_batched, const float *R, int R_batched, const float *t,
                  int t_batched, const float *dist_coeffs, int dist_batched, float orig_size, int B, int Vh, int Vo,
                  const float *grad_ndc, const float *grad_attrs, int has_ndc1, int has_ndc2, int has_attrs12,
                  int has_attrs21, float *grad_verts1, float *grad_verts2, void *stream);
/* hoc_flow_finalize: everything get_opticalflow does after its two renders (opticalflow.py:109-154): alpha
 * threshold, ignore-face mask, flow = rgb * mask, forward-backward occlusion check, mask products, channel
 * slice, crop.  rgb [B,3,S,S] / alpha [B,S,S] in HOC_LAYOUT_IMAGE, idx [B,S,S] raster order;
 * ignore_faces: device int[n_ignore] (n_ignore <= 64);  outputs flow12 / flow21 [B,H,W,2] and
 * mult1 / mult2 [B,H,W] = d flow / d rgb (kept for the backward). */
int hoc_flow_finalize(const float *rgb1, const float *alpha1, const int32_t *idx1, const float *rgb2,
                      const float *alpha2, const int32_t *idx2, int B, int S, int H, int W, const int *ignore_faces,
                      int n_ignore, int mask_occlusions, float distance_thresh, float *flow12, float *flow21,
                      float *mult1, float *mult2, void *stream);
/* grad_flow [B,H,W,2], mult [B,H,W] -> grad_rgb [B,3,S,S] (fully overwritten; zero outside the crop). */
int hoc_flow_finalize_backward(const float *grad_flow, const float *mult, int B, int S, int H, int W,
                               float *grad_rgb, void *stream);
/* Both directions of a pair in one launch. */
int hoc_flow_finalize_backward_pair(const float *grad_flow12, const float *mult1, const float *grad_flow21,
                                    const float *mult2, int B, int S, int H, int W, float *grad_rgb1, float *grad_rgb2,
                                    void *stream);

/* Per-vertex front end of get_opticalflow for one frame pair: batch_proj2d of both frames, the displacement
 * attributes [dx, dy, 1] of both directions (opticalflow.py:98-102,121-122) and nr.projection of both meshes
 * to NDC (renderer.py:187; OpenCV distortion, y flip, [-1,1] scaling by orig_size) in ONE launch.
 * verts1/verts2 [B,V,3]; K1/K2 [B or 1,3,3]; R [B or 1,3,3]; t [B or 1,3]; dist_coeffs [B or 1,5]
 * (`*_batched` = 1 when the leading dimension is B);  outputs ndc1/ndc2/attrs12/attrs21 [B,V,3]. */
int hoc_flow_vertices(const float *verts1, const float *verts2, const float *K1, int K1_batched, const float *K2,
                      int K2_batched, const float *R, int R_batched, const float *t, int t_batched,
                      const float *dist_coeffs, int dist_batched, float orig_size, int B, int V, float *ndc1,
                      float *ndc2, float *attrs12, float *attrs21, void *stream);
/* Adjoint: gradients of the four outputs (any may be NULL = zero) -> grad_verts1 / grad_verts2 [B,V,3]
 * (either may be NULL; fully overwritten).  Cameras get no gradient. */
int hoc_flow_vertices_backward(const float *verts1, const float *verts2, const float *K1, int K1_batched,
                               const float *K2, int K2_batched, const float *R, int R_batched, const float *t,
                               int t_batched, const float *dist_coeffs, int dist_batched, float orig_size, int B, int V,
                               const float *grad_ndc1, const float *grad_ndc2, const float *grad_attrs12,
                               const float *grad_attrs21, float *grad_verts1, float *grad_verts2, void *stream);

/* ---- MANO linear-blend skinning ------------------------------------------------------------
 * Replaces manopth ManoLayer.forward (absent dependency; called at manobranch.py:70-85,139-145 and built at
 * warpreg.py:54-60) for PCA or axis-angle pose input with an axis-angle root.  One launch per direction.
 * Model constants (device, fp32, MANO's own layouts): v_template [V,3], shapedirs [V,3,10],
 * posedirs [V,3,135], j_regressor [16,V], weights [V,16], hands_components [ncomps,45] (PCA rows actually
 * used; ignored when use_pca == 0), hands_mean [45] (zeros for flat_hand_mean).
 *   pose  [B,3+ncomps] (use_pca) or [B,48] (axis-angle, ncomps = 45); betas [B,10] or NULL (zeros);
 *   trans [B,3] or NULL (then the outputs are centred on reordered joint `center_idx`, -1 = no centring);
 *   verts [B,V,3], joints [B,21,3] out, millimetres and joint order of manopth.
 * weights [V,16] must be 16-byte aligned. */
typedef struct hoc_mano_model {
    const float *v_template;  /* [V,3] */
    const float *shapedirs;   /* [V,3,10] */
    const float *posedirs;    /* [V,3,135] */
    const float *j_regressor; /* [16,V] (kept for reference; the kernels use the two folded constants below) */
    /* derived constants, computed once per model by the caller:
     *   posedirs_t  [135][3V]    = posedirs transposed (coalesced per-coordinate reads)
     *   j_template  [16][3]      = j_regressor . v_template
     *   j_shapedirs [16][3][10]  = j_regressor . shapedirs   (J = j_template + j_shapedirs . betas) */
    const float *posedirs_t;
    const float *j_template;
    const float *j_shapedirs;
    const float *weights;
    const float *hands_components;
    const float *hands_mean;
    int num_verts;
    int ncomps;
    int use_pca;
    int center_idx;
    int tip_ids[5]; /* fingertip vertex ids (manopth: 745, 317, 444 (right) / 445 (left), 556, 673) */
} hoc_mano_model;

int hoc_mano_forward(const hoc_mano_model *model, const float *pose, const float *betas, const float *trans, int B,
                     float *verts, float *joints, void *stream);
/* grad_verts [B,V,3] / grad_joints [B,21,3] (either may be NULL) -> grad_pose [B,3+ncomps], grad_betas [B,10],
 * grad_trans [B,3] (any may be NULL; fully overwritten).  workspace: hoc_mano_backward_workspace_bytes(B) bytes
 * (per-sample accumulators of the vertex-parallel pass). */
size_t hoc_mano_backward_workspace_bytes(int B);
int hoc_mano_backward(const hoc_mano_model *model, const float *pose, const float *betas, const float *trans,
                      const float *grad_verts, const float *grad_joints, int B, float *grad_pose, float *grad_betas,
                      float *grad_trans, void *workspace, size_t workspace_bytes, void *stream);

/* ---- geometry head: network outputs -> camera-space meshes ---------------
 * Shared camera arguments (recover_3d_proj, /root/reference/meshreg/models/project.py:5-23):
 *   camintr [B,3,3] (or [1,3,3] with camintr_batched = 0), scale [B], trans [B,2] = the predicted pixel-space
 *   scale / translation BEFORE the factors (meshregnet.py:220-221, objbranch.py:50-51), scale_factor / trans_factor,
 *   off_z (0.4), input_res (res_w, res_h).
 *     Z0 = camintr[b,0,0] * scale * scale_factor + off_z
 *     XY0 = (trans * trans_factor + input_res / 2 - camintr[b,:2,2]) * Z0 / camintr[b,0,0];  center3d = (XY0, Z0)
 *
 * hoc_hand_head_forward replaces the geometry of MeshRegNet.recover_mano (meshregnet.py:191-229): ManoAdaptor
 * (bias-free Linear 778 -> 21, `adaptor` [J,V]; meshregnet.py:23-51), centring on adapted joint `center_idx`,
 * recover_3d_proj, recov_joints3d / recov_handverts3d and both batch_proj2d.  One launch.
 *   verts [B,V,3] (V <= 1024);  adaptor [J,V] or NULL (then joints_in [B,J,3] are the joints and nothing is
 *   re-derived from the vertices);  center_idx -1 = no centring (the no-adaptor branch of the reference);
 *   outputs (any may be NULL): joints3d [B,J,3], verts3d [B,V,3] (centred), recov_joints3d, recov_verts3d
 *   (+ center3d), joints2d [B,J,2], verts2d [B,V,2] (pixels), center3d [B,3]. */
int hoc_hand_head_forward(const float *verts, const float *joints_in, const float *adaptor, int B, int V, int J,
                          int center_idx, const float *camintr, int camintr_batched, const float *scale,
                          const float *trans, float scale_factor, float trans_factor, float off_z, float res_w,
                          float res_h, float *joints3d, float *verts3d, float *recov_joints3d, float *recov_verts3d,
                          float *joints2d, float *verts2d, float *center3d, void *stream);
/* Adjoint.  recov_verts3d / recov_joints3d: the forward's outputs (needed when a 2-D gradient is given);
 * g_*: gradients of the seven outputs (any may be NULL = zero);  results (any may be NULL; fully overwritten):
 * grad_verts [B,V,3], grad_joints_in [B,J,3] (adaptor == NULL) or grad_adapt [B,J,3] (d L / d adapted joints, from
 * which the caller forms a weight gradient if the adaptor is not frozen), grad_scale [B], grad_trans [B,2]. */
int hoc_hand_head_backward(const float *recov_verts3d, const float *recov_joints3d, const float *adaptor, int B, int V,
                           int J, int center_idx, const float *camintr, int camintr_batched, const float *scale,
                           const float *trans, float scale_factor, float trans_factor, float off_z, float res_w,
                           float res_h, const float *g_joints3d, const float *g_verts3d,
                           const float *g_recov_joints3d, const float *g_recov_verts3d, const float *g_joints2d,
                           const float *g_verts2d, const float *g_center3d, float *grad_verts, float *grad_joints_in,
                           float *grad_adapt, float *grad_scale, float *grad_trans, void *stream);
/* hoc_recover_points_forward replaces ObjBranch.forward's geometry (objbranch.py:46-77): batch_rodrigues of
 * rotaxisang [B,3] (manopth rodrigues_layer; NULL = no rotation, which makes the call recover_3d_proj itself),
 * rot_points = R . points, recov_points = rot_points + center3d, points2d = batch_proj2d(recov_points).
 * points [B,N,3]; outputs (any may be NULL): rot_points [B,N,3], recov_points [B,N,3], points2d [B,N,2], center3d [B,3]. */
int hoc_recover_points_forward(const float *points, const float *rotaxisang, int B, int N, const float *camintr,
                               int camintr_batched, const float *scale, const float *trans, float scale_factor,
                               float trans_factor, float off_z, float res_w, float res_h, float *rot_points,
                               float *recov_points, float *points2d, float *center3d, void *stream);
/* Adjoint: g_* any may be NULL;  grad_points [B,N,3], grad_rot [B,3], grad_scale [B], grad_trans [B,2]
 * (any may be NULL; fully overwritten).  Deterministic (sums are reduced inside one CTA per sample). */
int hoc_recover_points_backward(const float *points, const float *rotaxisang, int B, int N, const float *camintr,
                                int camintr_batched, const float *scale, const float *trans, float scale_factor,
                                float trans_factor, float off_z, float res_w, float res_h, const float *g_rot_points,
                                const float *g_recov_points, const float *g_points2d, const float *g_center3d,
                                float *grad_points, float *grad_rot, float *grad_scale, float *grad_trans,
                                void *stream);

/* ---- frame-pair input pipeline ---------------------------------------------------------------
 * The image side of HandObjSet.get_sample for the two frames of a pair
 * (/root/reference/meshreg/datasets/handobjset.py:336-379: colortrans.apply_jitter, handutils.transform_img, to_tensor,
 * normalize, the jitter mask), bit-compatible with the PIL / torchvision calls those helpers make (csrc/input_pipe.cu).
 *   frame0 / frame1  [B,Hs,Ws,3] uint8, the decoded source frames (HWC, as PIL holds them)
 *   coef_fix16       [B,6] int32: PIL's 16.16 fixed-point coefficients of the output -> input affine map (shared by the
 *                    two frames of a sample: the same space augmentation, handobjset.py:417-421): with
 *                    (a, b, c, d, e, f) = inv(affinetrans)[:2] they are floor(v * 65536 + 0.5) of
 *                    a, b, c + a / 2 + b / 2, d, e, f + d / 2 + e / 2
 *   color            [B,3] float: brightness, saturation, contrast factors;  hue_shift [B] int32: uint8(hue * 255)
 *   order            [B,2,4] int32: per (sample, frame) the order of the adjustments (0 brightness, 1 saturation, 2 hue,
 *                    3 contrast, -1 none); NULL = no colour jitter (evaluation mode)
 *   image0 / image1  [B,3,H,W] float out: x / 255 - 0.5;  mask0 / mask1 [B,3,H,W] float out (or both NULL): 1 where the
 *                    source pixel exists
 *   workspace        hoc_augment_frame_pair_workspace_bytes(B) bytes (the grey sums of the contrast adjustment)
 * Two launches (one without colour jitter). */
size_t hoc_augment_frame_pair_workspace_bytes(int B);
int hoc_augment_frame_pair(const uint8_t *frame0, const uint8_t *frame1, int B, int Hs, int Ws, const int *coef_fix16,
                           const float *color, const int *hue_shift, const int *order, int H, int W, float *image0,
                           float *image1, float *mask0, float *mask1, void *workspace, size_t workspace_bytes,
                           void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HOC_B200_H */
