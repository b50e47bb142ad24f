/*
 * read_b200 — C ABI of the B200-native READ render hot path.
 *
 * Plain C: pointers, sizes, ints.  No torch / C++ types cross this boundary.
 * All data pointers are CALLER-OWNED DEVICE memory unless a parameter is named
 * `*_host`.  Every launch is asynchronous on the caller's stream (`stream` is a
 * cudaStream_t passed as void*); no entry point synchronises the device,
 * allocates per call, or touches the host copy of the data.
 *
 * Return value: 0 on success, negative error code otherwise;
 * read_last_error() returns a thread-local message (the Python host raises
 * RuntimeError with it, matching the reference's AT_ASSERTM -> RuntimeError
 * behaviour, pcpr_cuda.cpp:17-21).
 *
 * Which reference interface each entry point replaces is cited per function
 * (paths relative to the JOP-Lee/READ tree).
 */
#ifndef READ_B200_H
#define READ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define READ_MAX_LEVELS 8
#define READ_MAX_SRC 4

enum {
    READ_OK = 0,
    READ_ERR_INVALID = -1,   /* bad argument (shape, alignment, null)            */
    READ_ERR_CUDA = -2,      /* CUDA runtime / driver error, see read_last_error */
    READ_ERR_UNSUPPORTED = -3
};

int read_version(void);
const char *read_last_error(void);
/* Tuning options.  Results are bit-identical for every accepted setting (one exception: "tc_pair_wide" selects a kernel whose K order
 * differs - within one bf16 ulp per layer output); unknown names and out-of-range values are rejected.
 *   rasterizer: "raster_pipelined" (1), "raster_bulk_tma" (1), "raster_mode" (0..3, default 2), "raster_occupancy" (0 = auto), "raster_stream" (1),
 *               "raster_dedup" (0), "raster_run" (0 = auto), "raster_nbr_filter" (0), "raster_stages" (2; 3 = deeper point ring),
 *               "raster_carveout" (45: preferred shared-memory carveout in percent for the streaming kernel, -1 = driver default)
 *   convs:      read when a plan is created: "tc_mt" (supertile width 1 (default) / 2 / 4, 0 = auto-widen), "tc_merge_done" (1),
 *               "tc_commit_late" (0), "tc_bpair" (0), "tc_probe" (0), "tc_pair" (1: CTA-pair cta_group::2 kernel for the Cin 64 layers and
 *               the Cin 32 layers without a residual, 2: every eligible layer, 0: off), "tc_pair_wide" (1: streamed-weight CTA-pair kernel for
 *               the Cin, Cout = 128 / 256 layers), "tc_wide_ntile" (256; 128 = narrower units for Cout 256, measured slower),
 *               "tc_tma_store" (1: epilogue items staged in shared memory and written by TMA stores; also read at launch by the pair kernel);
 *               read at launch: "tc_role_rot" (1), "tc_pdl" (1: programmatic dependent launch between consecutive conv kernels)
 *   The Python host applies READ_B200_OPTIONS="name=value,..." from the environment when it loads the library.
 * The options are process-wide tuning state (plain ints): set them before creating plans / launching, not concurrently with
 * launches from other threads.  Diagnostic knobs that skip work and therefore corrupt the output ("tc_debug", "tcg_debug",
 * raster_mode 4 / 5) exist only in builds compiled with -DREAD_DIAG and are absent from the shipped library. */
int read_set_option(const char *name, int value);
/* 1 if the current device is sm_100 (B200); the library refuses to launch elsewhere. */
int read_device_ok(void);

/* ------------------------------------------------------------------------------------------
 * Rasterizer.  Packed z-buffer entry: (uint64)float_bits(depth) << 32 | point_id ;
 * empty = 0x7FFFFFFFFFFFFFFF (max int64, so signed and unsigned min agree).  atomicMin on it == "min depth, ties -> lowest id", the
 * sequential semantics of DepthProject (MyRender/CloudProjection/point_render.cu:125-167).
 * Pyramid layout: level-major; level l holds [B, h_l, w_l] entries, w_l = int(W*0.5^l),
 * h_l = int(H*0.5^l) (src/READ/gl/myrender.py:33-34).
 * ---------------------------------------------------------------------------------------- */

/* Number of uint64 entries of a B-view, L-level pyramid (size your zbuf with this). */
int64_t read_pyramid_entries(int B, int W, int H, int L);
/* Entry offset of level l inside the pyramid, and its (w,h). */
int64_t read_pyramid_level_offset(int B, int W, int H, int l);
void read_level_size(int W, int H, int l, int *w, int *h);

/* Reset a pyramid (or any zbuf) to "empty". */
int read_zbuf_clear(uint64_t *zbuf, int64_t entries, void *stream);

/*
 * Replaces: the L calls of pcpr.forward made by MyRender.render
 * (src/READ/gl/myrender.py:32-40 -> pcpr_cuda.cpp:23-37 -> point_render.cu:169-200),
 * fused: every point is read, culled and projected ONCE for all B views and all L levels.
 *   xyz      [n,3] f32 device            total_m [B,16] f32 device, row-major 4x4 (proj @ inv(view))
 *   zbuf     pyramid of read_pyramid_entries(B,W,H,L) uint64, already cleared
 *   id_base  added to the local point index to form the global id written in the z-buffer
 *            (point-sharded multi-GPU rendering: each rank passes its shard's first global id)
 * Levels whose size is exactly half of the previous one are derived from it by a 2x2 min
 * (bit-identical to rasterising them directly); the others get direct atomics.
 */
int read_raster_project(const float *xyz, int64_t n, int64_t id_base, const float *total_m, int B,
                        int W, int H, int L, uint64_t *zbuf, void *stream);
/* Second half of the above when a collective sits in between (multi-GPU): project only the levels
 * that need direct atomics (read_raster_project_direct), all-reduce(min) them, then derive the rest. */
int read_raster_project_direct(const float *xyz, int64_t n, int64_t id_base, const float *total_m, int B,
                               int W, int H, int L, uint64_t *zbuf, void *stream);
int read_raster_derive_levels(int B, int W, int H, int L, uint64_t *zbuf, void *stream);
/* Single-view frame path over a spatially sorted point store: pts4 is [n,4] f32 = (x, y, z, bit pattern of the ORIGINAL
 * point id), ordered so that neighbouring points are neighbours in space (read_b200.ops.SortedPoints sorts by the Morton
 * code of the 3-D grid cell when the scene is loaded, the equivalent of MyRender.update_ds).  Rasterises level 0 only
 * (nested levels; finish with read_raster_derive_levels or read_pyramid_resolve_gather).  The z-buffer is a min over
 * (depth | original id) keys, hence identical to read_raster_project_direct on the unsorted cloud. */
int read_raster_project_sorted(const float *pts4, int64_t n, const float *total_m, int W, int H, int L, uint64_t *zbuf,
                               void *stream);
/* Same for B <= 8 views in ONE pass over the store (total_m [B,16]; view b goes to level 0 of view b of a B-view pyramid,
 * i.e. zbuf + b*W*H): the multi-GPU frame path, where every rank rasterises its spatial tile for all views of the step. */
int read_raster_project_sorted_views(const float *pts4, int64_t n, const float *total_m, int B, int W, int H, int L,
                                     uint64_t *zbuf, void *stream);
/* Bitmask of levels rasterised with direct atomics (bit l set) for this geometry. */
unsigned read_raster_direct_mask(int W, int H, int L);

/*
 * Replaces: the float outputs of GPU_PCPR (point_render.cu:176-177,196-199): for one level,
 * index [B,h,w] f32 (0 = empty) and depth [B,h,w] f32 (0 = empty) from the packed z-buffer.
 * Either output may be NULL.
 */
int read_zbuf_resolve(const uint64_t *zbuf_level, int64_t pixels, float *index_out, float *depth_out,
                      void *stream);

/*
 * Replaces: pcpr.forward itself (pcpr_cuda.cpp:23-37): one level, B views.
 * zbuf_ws: workspace of B*h*w uint64 (cleared inside).  Outputs as GPU_PCPR.
 */
int read_pcpr_forward(const float *xyz, int64_t n, const float *total_m, int B, int w, int h,
                      uint64_t *zbuf_ws, float *index_out, float *depth_out, void *stream);

/* ------------------------------------------------------------------------------------------
 * Descriptor gather (PointTexture, READ/models/texture.py:42-70).
 * Descriptors live point-major [N, D] ("shadow" of the checkpoint's [1, D, N] parameter) so that
 * one pixel touches one 32-byte sector (D = 8 f32) instead of 8.
 * ---------------------------------------------------------------------------------------- */

/* [1,D,N] f32 channel-major  <->  [N,D] f32 point-major. */
int read_texture_to_point_major(const float *tex_cn, int D, int64_t N, float *tex_nd, void *stream);
int read_texture_to_channel_major(const float *tex_nd, int D, int64_t N, float *tex_cn, void *stream);

enum { READ_FEAT_NCHW_F32 = 0, READ_FEAT_NHWC_F32 = 1, READ_FEAT_NHWC_BF16 = 2 };
enum { READ_TEXACT_NONE = 0, READ_TEXACT_SIGMOID = 1, READ_TEXACT_TANH = 2 };

/* From a float index map (API path: PointTexture.forward(ids), texture.py:52-63).
 * ids [pixels] f32 (channel 0 of the uv input, already contiguous); out [B, D, h, w] or NHWC. */
int read_gather_from_index(const float *tex_nd, int D, int64_t N, const float *ids, int B, int h, int w,
                           int layout, int activation, void *out, void *stream);
/* Fused path: straight from the packed z-buffer level (skips the float index map). */
int read_gather_from_zbuf(const float *tex_nd, int D, int64_t N, const uint64_t *zbuf_level, int B, int h,
                          int w, int layout, int activation, void *out, void *stream);
/* Per-frame fast path for a 4-level NESTED pyramid (every level an exact half of the previous one, W,H % 8 == 0)
 * and D == 8: one kernel derives levels 1..3 from level 0 (bit-identical 2x2 min), stores them, gathers all four
 * feature maps (NHWC bf16 or f32; outs = HOST array of 4 device pointers) and, if reset_level0, leaves level 0
 * cleared for the next frame.  Views [view0, view0+nviews) of the B-view pyramid are processed (outs hold nviews views).
 * Call after read_raster_project_direct (replaces derive + 4 gathers + next clear). */
int read_pyramid_resolve_gather(const float *tex_nd, int D, int64_t N, uint64_t *zbuf, int B, int view0, int nviews,
                                int W, int H, int L, int layout, void *const *outs, int reset_level0, void *stream);
/* Backward of the gather (autograd of index_select, texture.py:61): grad_tex_nd[ids[p], :] += grad_out[.., p]
 * grad_out is NCHW f32 [B,D,h,w]; grad_tex_nd [N,D] f32 accumulates (caller zeroes).  Empty pixels carry
 * id 0, so point 0 receives their gradient exactly like the reference. */
int read_gather_backward(const float *grad_out, const float *ids, int B, int D, int h, int w, int64_t N,
                         float *grad_tex_nd, void *stream);

/* ------------------------------------------------------------------------------------------
 * Gated convolution (BasicConv, READ/models/unet.py:22-53) with everything around it fused:
 *   y = bn_scale * ( A(conv_f(x)+b_f) * sigmoid(conv_m(x)+b_m) ) + bn_shift  [+ residual]
 * x is a VIRTUAL concat of up to 4 NHWC sources, each resampled on the fly (nearest up/down by an
 * integer factor = F.interpolate default mode, unet.py:239-250; bilinear x4 align_corners=False =
 * nn.Upsample, unet.py:200), optionally multiplied elementwise by `mul` (FAM, unet.py:114-117).
 * ---------------------------------------------------------------------------------------- */
enum { READ_ACT_F32 = 0, READ_ACT_BF16 = 1 };
enum { READ_SRC_IDENTITY = 0, READ_SRC_NEAREST_DOWN = 1, READ_SRC_NEAREST_UP = 2, READ_SRC_BILINEAR_UP4 = 3 };
enum { READ_OUT_NHWC = 0, READ_OUT_NCHW_F32 = 1,
       /* pre-activation accumulators [conv_f | conv_m] (no bias / activation / BN), NHWC with 2*Cout channels: one term of a
        * 1x1 conv over a concat whose other sources live at a finer resolution (see `addin`) */
       READ_OUT_RAW_NHWC = 2 };
enum { READ_CONV_AUTO = 0, READ_CONV_GENERIC = 1, READ_CONV_TCGEN05 = 2, READ_CONV_TCGEN05_GATHER = 3 };

typedef struct read_src {
    const void *ptr;   /* [B, H, W, C] NHWC, activation dtype */
    int32_t C, H, W;
    int32_t mode;      /* READ_SRC_* */
    int32_t factor;    /* resample factor for NEAREST_* (2,4,8); ignored otherwise */
} read_src;

typedef struct read_conv_desc {
    int32_t act_dtype;              /* READ_ACT_* : storage type of sources / residual / NHWC out */
    int32_t n_src;
    read_src src[READ_MAX_SRC];
    const void *mul;                /* optional [B,Hin,Win,Cin], only with n_src==1 identity */
    int32_t B, Hin, Win, Cin;       /* logical (post-resample, post-concat) input */
    int32_t Hout, Wout, Cout;
    int32_t k, stride, pad;
    int32_t elu;                    /* 1: A = ELU(alpha=1); 0: identity */
    const float *w_generic;         /* packed f32 [k*k*Cin][Npad] (read_pack_weights_generic) or NULL */
    const void *w_tc;               /* packed bf16 for the tcgen05 kernel (read_pack_weights_tc) or NULL */
    const float *bias_f, *bias_m;   /* [Cout] */
    const float *bn_scale, *bn_shift; /* [Cout] folded eval-mode BatchNorm */
    const void *residual;           /* optional [B,Hout,Wout,Cout] added after BN (ResBlock / FAM skip) */
    void *out;
    int32_t out_mode;               /* READ_OUT_* */
    void *out2;                     /* optional second NHWC output: out2 = y * out2_mul (feeds a FAM) */
    const void *out2_mul;
    int32_t impl;                   /* READ_CONV_* : which kernel (must match the packed weights) */
    /* optional RAW tensor [B, ceil(Hout/2), ceil(Wout/2), 2*Cout] (activation dtype) added, nearest-upsampled x2, to the
     * accumulators BEFORE bias / activation.  A 1x1 conv commutes with nearest upsampling, so
     *   conv1x1(cat[a, up2(b)]) == conv1x1_a(a) + up2(conv1x1_b(b)):
     * the coarse sources of the AFF heads (unet.py:79-89,252-254) are convolved at their own resolution into RAW tensors
     * and enter here, instead of being gathered 4..64 times each at the fine resolution.  tcgen05 TMA kernel only. */
    const void *addin;
    int32_t addin_H, addin_W;
} read_conv_desc;

typedef struct read_conv_plan read_conv_plan;

/* Npad of the generic packing for a given Cout (multiple of 32, f|m halves per 32-channel group). */
int read_generic_npad(int Cout);
/* Pack torch-layout weights [Cout,Cin,k,k] f32 (device) into the kernel layouts (device). */
int read_pack_weights_generic(const float *wf, const float *wm, int Cout, int Cin, int k, float *out,
                              void *stream);
int64_t read_tc_weight_elems(int Cout, int Cin, int k);
/* K-chunking of the packed layout depends on the conv stride (32-channel chunks for stride 2): pack with the stride the
 * plan will be created with.  read_pack_weights_tc == stride 1. */
int read_pack_weights_tc_strided(const float *wf, const float *wm, int Cout, int Cin, int k, int stride, void *out_bf16,
                                 void *stream);
/* Same, geometry taken from the layer descriptor (stride AND the channel granularity of a virtual concat's sources):
 * the form a caller should use for any descriptor that read_conv_tc_supported accepts. */
int read_pack_weights_tc_for(const read_conv_desc *d, const float *wf, const float *wm, void *out_bf16, void *stream);
int read_pack_weights_tc(const float *wf, const float *wm, int Cout, int Cin, int k, void *out_bf16,
                         void *stream);
/* 1 if the tcgen05 TMA kernel supports this layer (stride-1 k x k / stride-2 3x3, 4x4 single source; 1x1 virtual concat of
 * identity / nearest-down sources), else 0. */
int read_conv_tc_supported(const read_conv_desc *d);
/* Same for the tcgen05 kernel with a gathered A operand (any stride / concat / resampling, bf16 activations);
 * it has its own weight packing. */
int read_conv_tcg_supported(const read_conv_desc *d);
int64_t read_tcg_weight_elems(int Cout, int Cin, int k);
int read_pack_weights_tcg(const float *wf, const float *wm, int Cout, int Cin, int k, void *out_bf16, void *stream);

/* Plan = validated descriptor + chosen kernel + TMA tensor maps.  Host-side only, no device work. */
int read_conv_plan_create(const read_conv_desc *d, read_conv_plan **out);
int read_conv_plan_launch(const read_conv_plan *p, void *stream);
int read_conv_plan_impl(const read_conv_plan *p);   /* READ_CONV_GENERIC / _TCGEN05 / _TCGEN05_GATHER */
/* Cap the persistent grid of a tensor-core plan at max_ctas CTAs (0 = one per SM, the default): a caller that runs two independent
 * layer chains on two streams gives each a share of the SMs so that both are resident at once (read_b200/engine.py). */
int read_conv_plan_set_max_ctas(read_conv_plan *p, int max_ctas);
/* Tile traversal order of a tcgen05 TMA plan: 0 = top-down (default), 1 = bottom-up.  Same result.  A layer that walks the image in the
 * opposite direction of its producer starts on the tiles the producer wrote LAST, i.e. the ones still in the 126 MB L2: consecutive
 * layers of a chain alternate (read_b200/engine.py). */
int read_conv_plan_set_tile_order(read_conv_plan *p, int reversed);
void read_conv_plan_destroy(read_conv_plan *p);

/* nn.Upsample(scale_factor=4, mode='bilinear') (unet.py:200), NHWC [B,h,w,C] -> [B,4h,4w,C], C % 8 == 0. */
int read_upsample_bilinear4(const void *in, int act_dtype, int B, int h, int w, int C, void *out, void *stream);

/* Layout converters at the net boundary. */
/* Viewer output path (replaces READ/gl/nn.py:123-124 `permute + cat alpha` and the flip of viewer.py:267):
 * RGB planes [3,H,W] f32 -> [H,W,4] f32 (alpha constant), optionally flipped vertically. */
int read_frame_to_rgba(const float *rgb_planes, int H, int W, int flip_vertical, float alpha, float *out_hwc4, void *stream);

/* Net-input staging for NetAndTexture's viewer options on the fused path (READ/models/compose.py:162-171): src = f32 NHWC
 * features [B,hs,ws,C] at render resolution; factor = supersampling (bilinear reduce exactly as F.interpolate(scale_factor=1/ss,
 * mode='bilinear')); last (nullable) = f32 [B,hs/factor,ws/factor,C] temporal-average state: out = (cur + last) / 2 when
 * have_last, and last := out; dst = NHWC in act_dtype (the engine's input buffer). */
int read_stage_net_inputs(const float *src, int B, int hs, int ws, int C, int factor, float *last, int have_last, int act_dtype,
                          void *dst, void *stream);

int read_nchw_f32_to_nhwc(const float *in, int B, int C, int H, int W, int act_dtype, void *out, void *stream);
int read_nhwc_to_nchw_f32(const void *in, int act_dtype, int B, int C, int H, int W, float *out, void *stream);

/* ------------------------------------------------------------------------------------------
 * Strip-parallel refinement net across GPUs: halo rows travel between neighbouring ranks through
 * peer-mapped mailboxes (cudaMalloc + CUDA IPC), one kernel per exchange, no NCCL on the path (csrc/halo.cu).
 * read_ipc_alloc: cudaMalloc `bytes` (zero-filled) and export its 64-byte IPC handle; read_ipc_open maps a peer's handle
 * (another process on the same node) with lazy peer access.  read_epoch_bump increments the device-resident frame counter.
 * read_halo_exchange: push `bytes` from src_up / src_dn into the neighbours' mailbox slots (peer pointers) and publish the
 * current epoch in their flag words; then wait until both neighbours' epochs arrived in this rank's flags and copy this rank's
 * mailbox slots into dst_top / dst_bot.  Null pointers = no neighbour on that side.  Asynchronous on `stream`. */
typedef struct read_halo_desc {
    const void *src_up, *src_dn;               /* local rows to send to the strip above / below */
    void *peer_up_slot, *peer_dn_slot;         /* destination slots inside the neighbours' mailboxes */
    void *peer_up_flag, *peer_dn_flag;         /* uint32 flag words inside the neighbours' mailboxes */
    const void *slot_from_up, *slot_from_dn;   /* this rank's mailbox slots (written by the neighbours) */
    const void *flag_from_up, *flag_from_dn;   /* this rank's flag words */
    void *dst_top, *dst_bot;                   /* halo rows of the local tensor */
    int64_t bytes;                             /* per direction, multiple of 16 */
    const void *epoch;                         /* uint32 device word, see read_epoch_bump */
    void *cta_counter;                         /* uint32 device word private to this exchange, zero-initialised */
} read_halo_desc;
int read_ipc_alloc(int64_t bytes, void **dev_ptr, unsigned char *handle64);
int read_ipc_open(const unsigned char *handle64, void **peer_ptr);
int read_ipc_close(void *peer_ptr);
int read_ipc_free(void *dev_ptr);
int read_epoch_bump(uint32_t *epoch, void *stream);
int read_halo_exchange(const read_halo_desc *d, void *stream);

/* ------------------------------------------------------------------------------------------
 * Descriptor side of the training step (replaces autograd's dense index_add_ of READ/models/texture.py:55-63
 * and the dense torch.optim.RMSprop of READ/pipelines/ogl.py:16,97-102).  grad_nd [N,D] f32 and touched [N] u8 are persistent,
 * zero-initialised accumulators owned by the caller.
 *   read_gather_backward_sparse : grad_nd[id,:] += grad_out[b,:,y,x] for every pixel (ids [B,h,w] f32, grad_out [B,D,h,w] f32);
 *                                 touched[id] = 1
 *   read_sparse_rmsprop_step    : for touched points only - square_avg (point-major [N,D]) decayed lazily by alpha^(step -
 *                                 last_step[i]), RMSprop update (momentum 0, not centered) applied to param_cn ([1,D,N], the
 *                                 checkpoint layout) AND to shadow_nd ([N,D], may be null); the point's grad row and flag are
 *                                 cleared.  step counts optimizer steps from 1.
 *   read_square_avg_dense       : the dense optimizer's square_avg [1,D,N] after `step` steps (state_dict export)
 *   read_compact_touched        : touched rows -> (id, grad[D]) pairs; *count (zeroed by the caller) receives their number
 *   read_scatter_pairs          : grad_nd[id,:] += grads[k,:], touched[id] = 1 (pairs received from other ranks) */
int read_gather_backward_sparse(const float *grad_out, const float *ids, int B, int D, int h, int w, int64_t N,
                                float *grad_nd, unsigned char *touched, void *stream);
int read_sparse_rmsprop_step(float *param_cn, float *shadow_nd, float *grad_nd, unsigned char *touched, float *square_avg,
                             int32_t *last_step, int64_t N, int D, int step, float lr, float alpha, float eps, float weight_decay,
                             void *stream);
int read_square_avg_dense(const float *square_avg, const int32_t *last_step, int64_t N, int D, int step, float alpha, float *out_cn,
                          void *stream);
int read_compact_touched(const float *grad_nd, const unsigned char *touched, int64_t N, int D, int32_t *count, int capacity,
                         int32_t *out_ids, float *out_grads, void *stream);
int read_scatter_pairs(const int32_t *ids, const float *grads, int n, int D, int64_t N, float *grad_nd, unsigned char *touched,
                       void *stream);

/* Counts kernels launched by this library since load (bench.py's gpu_launches claim). */
int64_t read_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* READ_B200_H */
