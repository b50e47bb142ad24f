// General gated convolution on the tcgen05 tensor cores: implicit GEMM whose A operand is GATHERED by four
// producer warps (cp.async 16-byte chunks, zero-fill for padding) directly into the 128B-swizzled K-major
// layout the UMMA descriptor expects.  Covers every BasicConv shape of READ/models/unet.py that the pure-TMA kernel
// (conv_tc.cu) cannot express as shifted tile loads:
//   * stride-2 3x3 and 4x4 convolutions (feat_extract[1,2,6], [3,4,7])
//   * virtual concat of up to 4 sources (SCM / AFF / decoder merges: unet.py:88,105,263,271,279)
//   * nearest up/down resampling of a source (F.interpolate, unet.py:239-250) and bilinear x4 (nn.Upsample, :200)
//   * channel counts that are only multiples of 8 (8->32 first conv, 32->3 last conv, SCM 16/56/120/248)
//   * NCHW fp32 output for the final layer
//
// GEMM view: M = 128 output pixels (8 rows x 16 cols), K = (tap, cin) flattened in 8-channel chunks and padded to
// 64-element blocks, N = conv_f | conv_m channels (<= 256 per tile).  Weights arrive by TMA; accumulators live in
// TMEM (double buffered); the epilogue is the same fused bias/ELU/sigmoid/BN/residual tail as conv_tc.cu.
//
// Warp roles (832 threads, 1 CTA/SM, persistent): warp0 = weight TMA + TMEM alloc, warp1 = MMA issuer,
// warps2-9 = A gather producers, one warp per ring stage (lane -> fixed 16-byte K chunk, 4 tile columns x 8 rows),
// warps10-25 = epilogue (TMEM lane quadrant = warp%4, four warps per quadrant interleaved over 8-column chunks).
// Stride is a template parameter, row validity replaces clamps, source offsets are 32-bit, the tile decode is
// division-free, accumulators form a TMEM ring.
#include "common.cuh"
#include "conv_common.cuh"
#include "ptx.cuh"
#include <cuda.h>
#include <mutex>
#include <new>

namespace rb {

#ifdef READ_DIAG
#define TCG_DBG(a_, bit_) (((a_).debug & (bit_)) != 0)
#else
#define TCG_DBG(a_, bit_) false
#endif

constexpr int G_THREADS = 832;
constexpr int G_EPI_WARP0 = 10;              // warps 10..25
constexpr int G_EPI_WARPS = 16;
constexpr int G_TW = 16, G_TH = 8;
constexpr int G_MAX_STAGES = 8;
constexpr uint32_t G_SMEM_BUDGET = 200 * 1024;
constexpr int G_TMEM_COLS = 512;
constexpr int G_MAX_ACC = 8;
constexpr int G_KBLK = 64;                     // elements per K block (128-byte rows, SWIZZLE_128B)

struct GSrc {
    const __nv_bfloat16 *ptr;
    int C, H, W, mode, shift, c_begin;   // shift = log2(resample factor)
};

// Per-source parameters the producers read from shared memory (dynamic indexing of the kernel-parameter constant bank
// is slow).  Nearest resampling is one formula for identity / down / up: src = (dst << shl) >> shr; a coordinate that
// is inside the (virtual) conv input is always inside the source, so no clamp is needed.
struct __align__(16) SrcS {
    const __nv_bfloat16 *ptr;
    int plane;                 // H*W*C elements per image (< 2^31)
    int C;
    int rs;                    // W*C (row stride in elements)
    int shl, shr, bil;
    int W, H, pad0_, pad1_;    // bilinear path only
};

struct GArgs {
    GSrc src[READ_MAX_SRC];
    int n_src;
    int B, Hin, Win, Cin, Hout, Wout, Cout, Cout_pad;
    int ksize, stride, pad;
    int K, kblocks;
    int n_tile, n_tiles;
    int tiles_x, tiles_y;
    float inv_tx, inv_ty, inv_nt;   // reciprocals for the division-free tile decode
    int stages;
    int nacc;                  // TMEM accumulator ring depth (each n_tile columns wide)
    int pdl;                   // launched with programmatic stream serialization
    int lean16;                // epilogue work items = (quadrant, 16-column chunk) dealt over tiles (NHWC, Cout 16 / 32 / 64)
    int debug;                 // diagnostic knobs ("tcg_debug"): 1 = epilogue only hand-shakes, 2 = no MMAs, 4 = no gather copies
    uint32_t a_bytes, b_bytes;
    int elu;
    const float *bias_f, *bias_m, *scale, *shift;
    const __nv_bfloat16 *residual;
    void *out;
    int out_mode;
    __nv_bfloat16 *out2;
    const __nv_bfloat16 *out2_mul;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival from this thread once all its prior cp.async have completed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t g_pack2(float a, float b)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}
// one F2FP per pair: lo -> bits [0,16) (lower address)
__device__ __forceinline__ uint32_t g_cvt2(float lo, float hi)
{
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ float2 g_unpack2(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&u)); }

// bilinear x4, align_corners=False (torch upsample_bilinear2d): src = max(0.25*(dst+0.5)-0.5, 0).  Only the fp32-parity
// engine and the unit tests route a bilinear source through this kernel (the bf16 engine upsamples with its own kernel).
__device__ __noinline__ void gather_bilinear_row(const SrcS &sv, const __nv_bfloat16 *sbase, int ix, int iy, bool valid,
                                                 uint32_t dst)
{
    uint4 o = make_uint4(0, 0, 0, 0);
    if (valid) {
        float fx = 0.25f * ((float)ix + 0.5f) - 0.5f;
        fx = fx < 0.f ? 0.f : fx;
        const int x0 = (int)fx;
        const int xp = (x0 < sv.W - 1) ? 1 : 0;
        const float lx = fx - (float)x0, hx = 1.f - lx;
        float fy = 0.25f * ((float)iy + 0.5f) - 0.5f;
        fy = fy < 0.f ? 0.f : fy;
        const int y0 = (int)fy;
        const int yp = (y0 < sv.H - 1) ? 1 : 0;
        const float ly = fy - (float)y0, hy = 1.f - ly;
        const __nv_bfloat16 *p = sbase + ((long long)y0 * sv.W + x0) * sv.C;
        const uint4 v00 = __ldg(reinterpret_cast<const uint4 *>(p));
        const uint4 v01 = __ldg(reinterpret_cast<const uint4 *>(p + (long long)xp * sv.C));
        const uint4 v10 = __ldg(reinterpret_cast<const uint4 *>(p + (long long)yp * sv.W * sv.C));
        const uint4 v11 = __ldg(reinterpret_cast<const uint4 *>(p + ((long long)yp * sv.W + xp) * sv.C));
        const uint32_t *a00 = &v00.x, *a01 = &v01.x, *a10 = &v10.x, *a11 = &v11.x;
        uint32_t r[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 f00 = g_unpack2(a00[q]), f01 = g_unpack2(a01[q]), f10 = g_unpack2(a10[q]), f11 = g_unpack2(a11[q]);
            r[q] = g_pack2(hy * (hx * f00.x + lx * f01.x) + ly * (hx * f10.x + lx * f11.x),
                           hy * (hx * f00.y + lx * f01.y) + ly * (hx * f10.y + lx * f11.y));
        }
        o = make_uint4(r[0], r[1], r[2], r[3]);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
}

template <int STRIDE>
__global__ void __launch_bounds__(G_THREADS, 1)
gated_conv_tc_gather_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (s_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem_al = smem_raw + (smem_base - s_u32(smem_raw));

    const uint32_t stage_bytes = a.a_bytes + a.b_bytes;
    const uint32_t ring_bytes = stage_bytes * (uint32_t)a.stages;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_al + ring_bytes);
    const uint32_t full0 = s_u32(bars);
    const uint32_t empty0 = full0 + 8 * G_MAX_STAGES;
    const uint32_t tfull0 = empty0 + 8 * G_MAX_STAGES;
    const uint32_t tempty0 = tfull0 + 8 * G_MAX_ACC;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + 2 * G_MAX_STAGES + 2 * G_MAX_ACC);
    float4 *s_par4 = reinterpret_cast<float4 *>(bars + 2 * G_MAX_STAGES + 2 * G_MAX_ACC + 2);   // Cout_pad x {bias_f, bias_m, scale, shift}

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int CP = a.Cout_pad;
    if (a.pdl) pdl_launch_dependents();
    // per-(K block, 16-byte chunk) decode table, built once per CTA so the producers' hot loop has no division:
    // {source index (-1 = zero padding of K), ky, kx, channel offset inside the source}
    int4 *s_tab = reinterpret_cast<int4 *>(s_par4 + CP);
    SrcS *s_src = reinterpret_cast<SrcS *>(s_tab + a.kblocks * 8);
    if (threadIdx.x < READ_MAX_SRC) {
        const GSrc &g = a.src[threadIdx.x < a.n_src ? threadIdx.x : 0];
        SrcS v;
        v.ptr = g.ptr;
        v.plane = g.H * g.W * g.C;
        v.C = g.C;
        v.rs = g.W * g.C;
        v.shl = g.mode == READ_SRC_NEAREST_DOWN ? g.shift : 0;
        v.shr = g.mode == READ_SRC_NEAREST_UP ? g.shift : 0;
        v.bil = g.mode == READ_SRC_BILINEAR_UP4 ? 1 : 0;
        v.W = g.W; v.H = g.H; v.pad0_ = 0; v.pad1_ = 0;
        s_src[threadIdx.x] = v;
    }
    for (int i = threadIdx.x; i < a.kblocks * 8; i += G_THREADS) {
        const int kel = i * 8;
        int4 e = make_int4(-1, 0, 0, 0);
        if (kel < a.K) {
            const int tap = kel / a.Cin;
            const int c = kel - tap * a.Cin;
            int si = 0;
            for (int q = 1; q < READ_MAX_SRC; ++q)
                if (q < a.n_src && c >= a.src[q].c_begin) si = q;
            e = make_int4(si, tap / a.ksize, tap % a.ksize, c - a.src[si].c_begin);
        }
        s_tab[i] = e;
    }
    for (int i = threadIdx.x; i < CP; i += G_THREADS)
        s_par4[i] = i < a.Cout ? make_float4(a.bias_f[i], 0.5f * a.bias_m[i], 0.5f * a.scale[i], a.shift[i])   // gate_folded
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == 1 && lane == 0) {
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(full0 + 8 * s, 32 + 1);   // the 32 lanes of the stage's gather warp + the weight-TMA thread
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int i = 0; i < G_MAX_ACC; ++i) {
            mbar_init(tfull0 + 8 * i, 1);
            mbar_init(tempty0 + 8 * i, a.lean16 ? (uint32_t)(a.n_tile >> 3) : (uint32_t)G_EPI_WARPS);   // arrivals per tile
        }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(s_u32(tmem_ptr_smem), G_TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // everything above touched only static data (weights' descriptors, folded constants): with programmatic dependent launch
    // it overlapped the previous kernel's tail; activations / residuals below need that kernel to have completed
    if (a.pdl) pdl_wait();

    const int m_tiles = a.tiles_x * a.tiles_y * a.B;
    const long long total_tiles = (long long)m_tiles * a.n_tiles;
    const int n_total = a.n_tile * a.n_tiles;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== weight TMA producer =====================
            uint32_t s = 0, ph = 0;
            for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int nt = (int)t - fdiv_small((int)t, a.inv_nt) * a.n_tiles;
                int row = nt * a.n_tile;
                for (int kb = 0; kb < a.kblocks; ++kb, row += n_total) {
                    mbar_wait(empty0 + 8 * s, ph ^ 1u);
                    const uint32_t fb = full0 + 8 * s;
                    mbar_arrive_expect_tx(fb, a.b_bytes);
                    tma_load_2d(&tmB, fb, smem_base + s * stage_bytes + a.a_bytes, 0, row);
                    if (++s == (uint32_t)a.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            // (Two issuers on alternate tiles were tried and dropped: both would walk the SAME stage ring, and an issuer
            // skipping the other's stages can get more than one mbarrier phase ahead of a slot - the parity wait then
            // aliases.  conv_tc.cu can do it because each of its issuers owns a half ring.)
            // The ~90-cycle latency of try_wait is taken off the per-stage chain by probing the NEXT stage's barrier
            // before the current stage's MMAs are issued.
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.n_tile >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t desc_hi = (uint32_t)(make_kmajor_desc(0, 1024u, 2u) >> 32);
            const uint32_t st16 = stage_bytes >> 4, ab16 = a.a_bytes >> 4;
            const uint32_t lo0 = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);
            uint32_t s = 0, ph = 0, lo = lo0;
            uint32_t acc = 0, acc_ph = 0;              // accumulator ring (see conv_tc.cu): tile i -> slot i % nacc
            bool ready = false;                        // the current stage's full barrier was already seen complete
            for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                mbar_wait(tempty0 + 8 * acc, acc_ph ^ 1u);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)a.n_tile;
                for (int kb = 0; kb < a.kblocks; ++kb) {
                    if (!ready) mbar_wait(full0 + 8 * s, ph);
                    uint32_t sn = s + 1, phn = ph;
                    if (sn == (uint32_t)a.stages) { sn = 0; phn ^= 1u; }
                    ready = mbar_try_wait(full0 + 8 * sn, phn);      // probe (the ring is continuous across tiles)
                    fence_proxy_async();          // gathered A rows were written through the generic proxy
                    tcgen05_fence_after();
#pragma unroll
                    for (int kk = 0; kk < G_KBLK / 16; ++kk) {
                        if (TCG_DBG(a, 2)) continue;
                        umma_bf16_lohi(d_tmem, lo + 2u * kk, lo + ab16 + 2u * kk, desc_hi, idesc,
                                       kk != 0 ? 1u : (kb != 0 ? 1u : 0u));
                    }
                    umma_commit(empty0 + 8 * s);
                    lo += st16;
                    s = sn; ph = phn;
                    if (s == 0) lo = lo0;
                }
                umma_commit(tfull0 + 8 * acc);
                if (++acc == (uint32_t)a.nacc) { acc = 0; acc_ph ^= 1u; }
            }
        }
    } else if (warp < G_EPI_WARP0) {
        // ===================== A gather producers: 8 warps, ONE WARP PER RING STAGE =====================
        // Warp w owns ring slot w (stages <= 8 warps; a slot must have ONE producer that sees every one of its phases, or
        // the parity wait on its empty barrier could alias two phases) and fills it on its own: lane -> (16-byte K chunk j = lane & 7, tile
        // column px = 4*g + (lane >> 3), g = 0..3), 8 tile rows each, 32 copies per lane and stage.  Round-1 knob
        // experiments: with 256 threads per stage the 256 serialised mbarrier arrivals alone cost ~380 cycles per
        // stage (3x the stage's MMA time) and every warp paid the wait / decode / table latencies of every stage; here
        // a stage costs 32 arrivals and its per-stage set-up is paid once per lane, while 8 stages are gathered
        // concurrently by the 8 warps.
        const uint32_t w = (uint32_t)(warp - 2);
        const uint32_t my_empty = empty0 + 8 * w;
        const int j = lane & 7;
        const int pq = lane >> 3;
        uint32_t s = 0, ph = 0;           // ring slot of the current iteration and its phase (tracked for EVERY iteration)
        const int4 *tab = s_tab + j;
        for (long long tl = blockIdx.x; tl < total_tiles; tl += gridDim.x) {
            const int mt = fdiv_small((int)tl, a.inv_nt);
            const int q_ = fdiv_small(mt, a.inv_tx);
            const int tx = mt - q_ * a.tiles_x;
            const int b = fdiv_small(q_, a.inv_ty);
            const int ty = q_ - b * a.tiles_y;
            const int oy0 = ty * G_TH;
            const int iyb = oy0 * STRIDE - a.pad;
            const int rows_ok = a.Hout - oy0;          // tile rows i < rows_ok produce output
            for (int kb = 0; kb < a.kblocks; ++kb) {
                if (s == w) {
                    mbar_wait(my_empty, ph ^ 1u);
                    const uint32_t dst0 = smem_base + s * stage_bytes;
                    if (TCG_DBG(a, 4)) {
                        mbar_arrive(full0 + 8 * s);
                    } else {
                        const int4 e = tab[kb * 8];
                        const SrcS &sv = s_src[e.x < 0 ? 0 : e.x];
                        const __nv_bfloat16 *sbase = sv.ptr + (long long)b * sv.plane;
                        const int iy0 = iyb + e.y;
                        if (!sv.bil) {
                            const int shl = sv.shl, shr = sv.shr, rs = sv.rs, C = sv.C;
                            // per-row source offsets (elements), -1 = row outside the image / the output
                            int roff[G_TH];
#pragma unroll
                            for (int i = 0; i < G_TH; ++i) {
                                const int iy = iy0 + i * STRIDE;
                                const bool rv = (e.x >= 0) && (i < rows_ok) && ((unsigned)iy < (unsigned)a.Hin);
                                roff[i] = rv ? ((iy << shl) >> shr) * rs + e.w : -1;
                            }
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const int px = g * 4 + pq;
                                const int ox = tx * G_TW + px;
                                const int ix = ox * STRIDE - a.pad + e.z;
                                const bool xv = (ox < a.Wout) && ((unsigned)ix < (unsigned)a.Win);
                                const int xoff = ((ix << shl) >> shr) * C;
                                const uint32_t dcol = dst0 + (uint32_t)px * 128u + (uint32_t)((j ^ (px & 7)) << 4);
#pragma unroll
                                for (int i = 0; i < G_TH; ++i) {
                                    const bool valid = xv && roff[i] >= 0;
                                    cp_async16(dcol + (uint32_t)i * 2048u, sbase + (valid ? roff[i] + xoff : 0), valid ? 16u : 0u);
                                }
                            }
                            // asynchronous arrival: the barrier counts this lane in when its copies above have landed (no
                            // wait, no writer-side proxy fence: the MMA thread fences once per stage after its wait)
                            cp_async_arrive_noinc(full0 + 8 * s);
                        } else {
                            for (int g = 0; g < 4; ++g) {
                                const int px = g * 4 + pq;
                                const int ox = tx * G_TW + px;
                                const int ix = ox * STRIDE - a.pad + e.z;
                                const bool xv = (e.x >= 0) && (ox < a.Wout) && ((unsigned)ix < (unsigned)a.Win);
                                const uint32_t dcol = dst0 + (uint32_t)px * 128u + (uint32_t)((j ^ (px & 7)) << 4);
                                for (int i = 0; i < G_TH; ++i) {
                                    const int iy = iy0 + i * STRIDE;
                                    const bool valid = xv && (i < rows_ok) && (iy >= 0) && (iy < a.Hin);
                                    gather_bilinear_row(sv, sbase + e.w, ix, iy, valid, dcol + (uint32_t)i * 2048u);
                                }
                            }
                            fence_proxy_async();          // plain st.shared data: writer-side fence + ordinary arrive
                            mbar_arrive(full0 + 8 * s);
                        }
                    }
                }
                if (++s == (uint32_t)a.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp < G_EPI_WARP0 + G_EPI_WARPS) {
        // ===================== epilogue (warps 10-25) =====================
        const int q = warp & 3;                       // TMEM lane quadrant this warp may read
        const int sub = (warp - G_EPI_WARP0) >> 2;    // 0..3: 8-column chunks sub, sub + 4, ...
        const int r = q * 32 + lane;
        const int py = r / G_TW, pxl = r % G_TW;
        const int half = a.n_tile >> 1;
        uint32_t acc_c = 0, acc_p = 0, lean_it = 0;
        for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int mt = fdiv_small((int)t, a.inv_nt);
            const int nt = (int)t - mt * a.n_tiles;
            const int q_ = fdiv_small(mt, a.inv_tx);
            const int tx = mt - q_ * a.tiles_x;
            const int b = fdiv_small(q_, a.inv_ty);
            const int ty = q_ - b * a.tiles_y;
            const int x = tx * G_TW + pxl, y = ty * G_TH + py;
            const bool inside = (x < a.Wout) && (y < a.Hout);
            const int pix = (b * a.Hout + y) * a.Wout + x;                  // output elements < 2^31 (checked by the host)
            const uint32_t acc = acc_c, acc_ph = acc_p;
            if (++acc_c == (uint32_t)a.nacc) { acc_c = 0; acc_p ^= 1u; }
            const bool nhwc = a.out_mode != READ_OUT_NCHW_F32;
            const bool has_res = a.residual != nullptr, has_out2 = a.out2 != nullptr;
            if (a.lean16) {
                // (quadrant, 16-column chunk) items, 4 * nch16 per tile, dealt warp -> chunk (sub % nch16) of every G-th
                // tile (G = 4 / nch16): a lane owns 32 contiguous output bytes (full sectors), see conv_tc.cu
                const int nch16 = half >> 4;
                const int chunk = sub & (nch16 - 1);
                if (((lean_it++) & (uint32_t)((4 >> (nch16 >> 1)) - 1)) != (uint32_t)(sub >> (nch16 >> 1))) continue;
                const int co = chunk * 16;
                const int o = pix * a.Cout + co;
                uint4 rs0 = make_uint4(0, 0, 0, 0), rs1 = rs0, ml0 = rs0, ml1 = rs0;
                if (inside) {
                    if (has_res) {
                        rs0 = __ldg(reinterpret_cast<const uint4 *>(a.residual + o));
                        rs1 = __ldg(reinterpret_cast<const uint4 *>(a.residual + o) + 1);
                    }
                    if (has_out2) {
                        ml0 = __ldg(reinterpret_cast<const uint4 *>(a.out2_mul + o));
                        ml1 = __ldg(reinterpret_cast<const uint4 *>(a.out2_mul + o) + 1);
                    }
                }
                mbar_wait(tfull0 + 8 * acc, acc_ph);
                tcgen05_fence_after();
                const uint32_t trow = tmem_base + acc * (uint32_t)a.n_tile + ((uint32_t)(q * 32) << 16);
                if (TCG_DBG(a, 1)) {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
                    continue;
                }
                // two 8-column halves in sequence (register budget: 72 per thread at 832 threads)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t f8[8], m8[8];
                    tmem_ld8(trow + (uint32_t)(co + 8 * hh), f8);
                    tmem_ld8(trow + (uint32_t)(half + co + 8 * hh), m8);
                    tmem_ld_wait();
                    if (hh == 1) {                    // accumulator fully read: hand the TMEM slot back before the math
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
                    }
                    float yv[8];
                    if (a.elu) {
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) yv[jj] = gate_folded<true>(__uint_as_float(f8[jj]), __uint_as_float(m8[jj]), s_par4[co + 8 * hh + jj]);
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) yv[jj] = gate_folded<false>(__uint_as_float(f8[jj]), __uint_as_float(m8[jj]), s_par4[co + 8 * hh + jj]);
                    }
                    if (inside) {
                        const uint4 rs = hh ? rs1 : rs0, ml = hh ? ml1 : ml0;
                        if (has_res) {
                            const uint32_t rr[4] = {rs.x, rs.y, rs.z, rs.w};
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                yv[2 * jj] += __uint_as_float(rr[jj] << 16);
                                yv[2 * jj + 1] += __uint_as_float(rr[jj] & 0xFFFF0000u);
                            }
                        }
                        uint32_t pk[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) pk[jj] = g_cvt2(yv[2 * jj], yv[2 * jj + 1]);
                        reinterpret_cast<uint4 *>(static_cast<__nv_bfloat16 *>(a.out) + o)[hh] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        if (has_out2) {
                            const uint32_t mm[4] = {ml.x, ml.y, ml.z, ml.w};
                            uint32_t p2[4];
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const float2 ys = g_unpack2(pk[jj]);
                                const float2 mv = g_unpack2(mm[jj]);
                                p2[jj] = g_cvt2(ys.x * mv.x, ys.y * mv.y);
                            }
                            reinterpret_cast<uint4 *>(a.out2 + o)[hh] = make_uint4(p2[0], p2[1], p2[2], p2[3]);
                        }
                    }
                }
                continue;
            }
            const int c_first = sub * 8;
            // first chunk's residual / FAM multiplier: issued before the wait for the MMAs
            uint4 r0 = make_uint4(0, 0, 0, 0), m0 = make_uint4(0, 0, 0, 0);
            if (inside && nhwc && c_first < half && nt * half + c_first < a.Cout) {
                const int o = pix * a.Cout + nt * half + c_first;
                if (has_res) r0 = __ldg(reinterpret_cast<const uint4 *>(a.residual + o));
                if (has_out2) m0 = __ldg(reinterpret_cast<const uint4 *>(a.out2_mul + o));
            }
            mbar_wait(tfull0 + 8 * acc, acc_ph);
            tcgen05_fence_after();
            if (TCG_DBG(a, 1)) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
                continue;
            }
            const uint32_t trow = tmem_base + acc * (uint32_t)a.n_tile + ((uint32_t)(q * 32) << 16);
            for (int c0 = c_first; c0 < half; c0 += 8 * (G_EPI_WARPS / 4)) {
                const int co = nt * half + c0;
                if (co >= a.Cout) break;             // padded channels (warp-uniform)
                uint32_t rf[8], rm[8];
                tmem_ld8(trow + (uint32_t)c0, rf);
                tmem_ld8(trow + (uint32_t)(half + c0), rm);
                const int o = pix * a.Cout + co;
                if (c0 != c_first && inside && nhwc) {
                    if (has_res) r0 = __ldg(reinterpret_cast<const uint4 *>(a.residual + o));
                    if (has_out2) m0 = __ldg(reinterpret_cast<const uint4 *>(a.out2_mul + o));
                }
                tmem_ld_wait();
                float yv[8];
                if (a.elu) {
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) yv[jj] = gate_folded<true>(__uint_as_float(rf[jj]), __uint_as_float(rm[jj]), s_par4[co + jj]);
                } else {
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) yv[jj] = gate_folded<false>(__uint_as_float(rf[jj]), __uint_as_float(rm[jj]), s_par4[co + jj]);
                }
                if (inside) {
                    if (!nhwc) {
                        float *op = static_cast<float *>(a.out);
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj)
                            if (co + jj < a.Cout) op[(((long long)b * a.Cout + co + jj) * a.Hout + y) * a.Wout + x] = yv[jj];
                    } else {
                        if (has_res) {
                            const uint32_t rr[4] = {r0.x, r0.y, r0.z, r0.w};
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                yv[2 * jj] += __uint_as_float(rr[jj] << 16);
                                yv[2 * jj + 1] += __uint_as_float(rr[jj] & 0xFFFF0000u);
                            }
                        }
                        uint32_t pk[4];
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) pk[jj] = g_cvt2(yv[2 * jj], yv[2 * jj + 1]);
                        *reinterpret_cast<uint4 *>(static_cast<__nv_bfloat16 *>(a.out) + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        if (has_out2) {
                            const uint32_t mm[4] = {m0.x, m0.y, m0.z, m0.w};
                            uint32_t p2[4];
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const float2 ys = g_unpack2(pk[jj]);
                                const float2 mv = g_unpack2(mm[jj]);
                                p2[jj] = g_cvt2(ys.x * mv.x, ys.y * mv.y);
                            }
                            *reinterpret_cast<uint4 *>(a.out2 + o) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, G_TMEM_COLS);
    }
}

// ------------------------------------------------------------------ weight packing
// out[((kb * n_total + n) * 64 + kk)],  k = kb*64 + kk = (tap*Cin + c) ; n -> (tile, f|m half, channel); zero padded
__global__ void pack_tcg_kernel(const float *__restrict__ wf, const float *__restrict__ wm, int Cout, int Cin, int k,
                                int kblocks, int n_tile, int n_tiles, __nv_bfloat16 *__restrict__ out)
{
    const int K = k * k * Cin;
    const int n_total = n_tile * n_tiles;
    const int half = n_tile / 2;
    const long long total = (long long)kblocks * n_total * G_KBLK;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(i % G_KBLK);
        long long r = i / G_KBLK;
        const int n = (int)(r % n_total);
        const int kb = (int)(r / n_total);
        const int kel = kb * G_KBLK + kk;
        const int nt = n / n_tile, rr = n % n_tile;
        const bool is_m = rr >= half;
        const int co = nt * half + (rr % half);
        float v = 0.f;
        if (kel < K && co < Cout) {
            const int tap = kel / Cin, c = kel % Cin;
            const int ky = tap / k, kx = tap % k;
            const float *w = is_m ? wm : wf;
            v = w[(((long long)co * Cin + c) * k + ky) * k + kx];
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------ host side
struct GGeom {
    int cout_pad, n_tile, n_tiles, kblocks;
};
static bool g_geom(int Cin, int Cout, int k, GGeom *g)
{
    if (Cin % 8 != 0 || Cout < 1) return false;
    int cp, n_tile, n_tiles;
    if (Cout <= 128) {
        cp = ((Cout + 7) / 8) * 8;
        n_tile = 2 * cp;
        n_tiles = 1;
    } else {
        cp = ((Cout + 127) / 128) * 128;
        n_tile = 256;
        n_tiles = cp / 128;
    }
    if (n_tile % 16 != 0 || n_tile > 256) return false;
    if (g) *g = GGeom{cp, n_tile, n_tiles, (k * k * Cin + G_KBLK - 1) / G_KBLK};
    return true;
}

bool tcg_supported(const read_conv_desc &d)
{
    if (d.act_dtype != READ_ACT_BF16 || d.mul != nullptr) return false;
    if (d.out_mode == READ_OUT_RAW_NHWC || d.addin != nullptr) return false;      // TMA kernel only
    if (d.out_mode == READ_OUT_NHWC && d.Cout % 8 != 0) return false;
    for (int i = 0; i < d.n_src; ++i) {
        if (d.src[i].C % 8 != 0) return false;
        const int f = d.src[i].factor;
        const bool resampled = d.src[i].mode == READ_SRC_NEAREST_DOWN || d.src[i].mode == READ_SRC_NEAREST_UP;
        if (resampled && (f < 2 || (f & (f - 1)) != 0)) return false;      // power-of-two factors only (shifts)
    }
    return g_geom(d.Cin, d.Cout, d.k, nullptr);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

struct TcgPlan {
    CUtensorMap tmB;
    GArgs args;
    size_t smem_bytes;
};

int tcg_plan_create(const read_conv_desc &d, TcgPlan **out)
{
    GGeom g;
    if (!tcg_supported(d) || !g_geom(d.Cin, d.Cout, d.k, &g)) {
        set_error("tcgen05 gather conv: unsupported layer");
        return READ_ERR_UNSUPPORTED;
    }
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) {
        set_error("tcgen05 gather conv: cuTensorMapEncodeTiled not available from the driver");
        return READ_ERR_CUDA;
    }
    RB_CHECK_ARG((reinterpret_cast<uintptr_t>(d.w_tc) & 127) == 0, "tcgen05 gather conv: packed weights must be 128B aligned");
    RB_CHECK_ARG((reinterpret_cast<uintptr_t>(d.out) & 15) == 0, "tcgen05 gather conv: output must be 16B aligned");
    TcgPlan *p = new (std::nothrow) TcgPlan{};
    RB_CHECK_ARG(p != nullptr, "tcgen05 gather conv: out of host memory");
    {
        const cuuint64_t rows = (cuuint64_t)g.kblocks * g.n_tile * g.n_tiles;
        cuuint64_t dims[2] = {(cuuint64_t)G_KBLK, rows};
        cuuint64_t strides[1] = {(cuuint64_t)G_KBLK * 2};
        cuuint32_t box[2] = {(cuuint32_t)G_KBLK, (cuuint32_t)g.n_tile};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&p->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(d.w_tc), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tcgen05 gather conv: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
            delete p;
            return READ_ERR_CUDA;
        }
    }
    GArgs &a = p->args;
    int cb = 0;
    for (int i = 0; i < d.n_src; ++i) {
        int sh = 0;
        while ((1 << sh) < d.src[i].factor) ++sh;
        a.src[i] = GSrc{static_cast<const __nv_bfloat16 *>(d.src[i].ptr), d.src[i].C, d.src[i].H, d.src[i].W, d.src[i].mode,
                        sh, cb};
        cb += d.src[i].C;
    }
    for (int i = d.n_src; i < READ_MAX_SRC; ++i) a.src[i] = a.src[0];
    a.n_src = d.n_src;
    a.B = d.B; a.Hin = d.Hin; a.Win = d.Win; a.Cin = d.Cin;
    a.Hout = d.Hout; a.Wout = d.Wout; a.Cout = d.Cout; a.Cout_pad = g.cout_pad;
    a.ksize = d.k; a.stride = d.stride; a.pad = d.pad;
    a.K = d.k * d.k * d.Cin; a.kblocks = g.kblocks;
    a.n_tile = g.n_tile; a.n_tiles = g.n_tiles;
    a.tiles_x = (d.Wout + G_TW - 1) / G_TW;
    a.tiles_y = (d.Hout + G_TH - 1) / G_TH;
    a.inv_tx = 1.0f / (float)a.tiles_x;
    a.inv_ty = 1.0f / (float)a.tiles_y;
    a.inv_nt = 1.0f / (float)a.n_tiles;
    a.nacc = G_TMEM_COLS / g.n_tile > G_MAX_ACC ? G_MAX_ACC : G_TMEM_COLS / g.n_tile;
    if ((long long)d.B * d.Hout * d.Wout * d.Cout >= (1ll << 31)) {
        set_error("tcgen05 gather conv: output too large for 32-bit offsets");
        delete p;
        return READ_ERR_UNSUPPORTED;
    }
    a.debug = 0;
    a.lean16 = (d.out_mode == READ_OUT_NHWC && g.n_tiles == 1 && g.cout_pad == d.Cout &&
                (d.Cout == 16 || d.Cout == 32 || d.Cout == 64)) ? 1 : 0;
    if ((long long)a.tiles_x * a.tiles_y * d.B * a.n_tiles >= (1ll << 22)) {
        set_error("tcgen05 gather conv: too many tiles for the division-free decode");
        delete p;
        return READ_ERR_UNSUPPORTED;
    }
    a.a_bytes = 128u * G_KBLK * 2u;
    a.b_bytes = (uint32_t)g.n_tile * G_KBLK * 2u;
    int stages = (int)(G_SMEM_BUDGET / (a.a_bytes + a.b_bytes));
    if (stages > G_MAX_STAGES) stages = G_MAX_STAGES;
    a.stages = stages;
    if (stages < 3 || (d.stride != 1 && d.stride != 2)) {
        set_error("tcgen05 gather conv: unsupported geometry (stages %d, stride %d)", stages, d.stride);
        delete p;
        return READ_ERR_UNSUPPORTED;
    }
    a.elu = d.elu;
    a.bias_f = d.bias_f; a.bias_m = d.bias_m; a.scale = d.bn_scale; a.shift = d.bn_shift;
    a.residual = static_cast<const __nv_bfloat16 *>(d.residual);
    a.out = d.out; a.out_mode = d.out_mode;
    a.out2 = static_cast<__nv_bfloat16 *>(d.out2);
    a.out2_mul = static_cast<const __nv_bfloat16 *>(d.out2_mul);
    p->smem_bytes = 1024 + (size_t)stages * (a.a_bytes + a.b_bytes) + 8 * (2 * G_MAX_STAGES + 2 * G_MAX_ACC + 2) + 16 * (size_t)g.cout_pad + 16 * (size_t)g.kblocks * 8 + sizeof(SrcS) * READ_MAX_SRC + 64;
    *out = p;
    return READ_OK;
}

int g_tcg_debug = 0;
extern int g_tc_pdl;

int tcg_plan_launch(const TcgPlan *p, cudaStream_t st, int max_ctas)
{
    GArgs a = p->args;
    a.debug = g_tcg_debug;
    const long long total_tiles = (long long)a.tiles_x * a.tiles_y * a.B * a.n_tiles;
    if (total_tiles == 0) return READ_OK;
    long long grid = num_sms();
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    if (grid > total_tiles) grid = total_tiles;
    a.pdl = g_tc_pdl ? 1 : 0;
    cudaLaunchAttribute lattr[1];
    lattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    lattr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t lcfg{};
    lcfg.gridDim = dim3((unsigned)grid);
    lcfg.blockDim = dim3(G_THREADS);
    lcfg.dynamicSmemBytes = p->smem_bytes;
    lcfg.stream = st;
    lcfg.attrs = lattr;
    lcfg.numAttrs = a.pdl ? 1 : 0;
    if (a.stride == 1) {
        RB_CUDA(cudaFuncSetAttribute(gated_conv_tc_gather_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
        RB_CUDA(cudaLaunchKernelEx(&lcfg, gated_conv_tc_gather_kernel<1>, p->tmB, a));
    } else {
        RB_CUDA(cudaFuncSetAttribute(gated_conv_tc_gather_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
        RB_CUDA(cudaLaunchKernelEx(&lcfg, gated_conv_tc_gather_kernel<2>, p->tmB, a));
    }
    RB_LAUNCH_CHECK();
    return READ_OK;
}

void tcg_plan_destroy(TcgPlan *p) { delete p; }

int64_t tcg_weight_elems(int Cout, int Cin, int k)
{
    GGeom g;
    if (!g_geom(Cin, Cout, k, &g)) return -1;
    return (int64_t)g.kblocks * g.n_tile * g.n_tiles * G_KBLK;
}

int tcg_pack(const float *wf, const float *wm, int Cout, int Cin, int k, void *out, cudaStream_t st)
{
    GGeom g;
    if (!g_geom(Cin, Cout, k, &g)) {
        set_error("pack_tc_gather: unsupported channel counts %d -> %d", Cin, Cout);
        return READ_ERR_INVALID;
    }
    const long long total = (long long)g.kblocks * g.n_tile * g.n_tiles * G_KBLK;
    long long blocks = (total + 255) / 256;
    if (blocks > 65535) blocks = 65535;
    pack_tcg_kernel<<<(unsigned)blocks, 256, 0, st>>>(wf, wm, Cout, Cin, k, g.kblocks, g.n_tile, g.n_tiles, (__nv_bfloat16 *)out);
    RB_LAUNCH_CHECK();
    return READ_OK;
}

}  // namespace rb
