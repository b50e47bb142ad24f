// CTA-PAIR variant of the tcgen05 gated convolution (conv_tc.cu) for the small-channel 3x3 layers: tcgen05.mma.cta_group::2.
//
// Why (profiles/r02_role_timelines.md): an SS-mode M=128 x N=64 x K=16 MMA reads 4 KB of A and 2 KB of B from shared memory for
// 32 tensor cycles of work - at 128 B/clk that is 48 cycles, so the C=32 layers (N = 64) run the tensor pipe at <= 67 % and
// the C=64 layers (N = 128: 8 KB per 64 cycles) exactly at the shared-memory limit; on top of that every 128-pixel tile costs
// the single issuing thread a fixed wait / issue / commit sequence.  A CTA pair (two SMs of one TPC, cluster of 2) issues ONE
// M=256 MMA for two 128-pixel tiles: each SM reads its own A tile (4 KB) but only HALF of B (the N dimension of the weights is
// split across the pair and shared by the hardware), the issue / commit / hand-shake sequence is paid once per TWO tiles, and
// only half of the weights are resident per SM (C=64: 72 KB instead of 144 KB).
//
// Scope: stride-1 k x k convs, one identity source, Cin == one K chunk (32 or 64), Cout in {16, 32, 64}, NHWC bf16 output with
// the lean epilogue (optional ELU / residual) - the ResBlock / FAM / AFF 3x3 layers of the full- and half-resolution stages
// (READ/models/unet.py:11-20,22-53).  Everything else stays on conv_tc.cu.
//
// Protocol (leader = cluster rank 0).  Barriers live at the same shared-memory offsets in both CTAs.
//   bres   (leader's)  both CTAs' resident-weight TMA loads signal it (cta_group::2 loads may target the peer's barrier)
//   afull  (leader's)  slot s: the leader's producer arms it with 2 x tile bytes, both CTAs' A loads complete on it
//   tfull  (each CTA)  slot s: ONE multicast tcgen05.commit per tile pair: "MMAs done" = accumulator ready for the epilogue AND
//                      A stage s free for both producers (A ring depth == accumulator ring depth, as conv_tc's merge_done)
//   tempty (leader's)  slot s: both CTAs' epilogue items arrive (remote arrive from the peer) once their TMEM loads are done
// Accumulators: cta_group::2 allocation, 128 lanes x N columns per tile in EACH CTA's TMEM at the same column offset.
// Issue: warps 1 and 3 of the leader take alternate units; epilogue: 16 warps per CTA, items staged in shared memory and written by
// TMA stores (residual tiles TMA-loaded one item ahead) - see the kernel.
//
// This file holds two kernels: gated_conv_tc2_kernel (weights resident: Cin 32 / 64) and, further down, gated_conv_tc2s_kernel
// (weights streamed: Cin, Cout = 128 / 256, where the pair halves the L2 -> SM weight traffic per pixel).
#include "common.cuh"
#include "conv_common.cuh"
#include "ptx.cuh"
#include <cuda.h>
#include <new>

namespace rb {

#ifdef READ_DIAG
// per-role timeline of cluster 0 (see conv_tc.cu TC_TRACE; same buffer / reader: scripts/tc_trace.py); role slot = rank * 8 + role
#define T2_TRACE(role_, code_)                                                                                          \
    do {                                                                                                                \
        if (a.trace != nullptr && (int)(role_) >= 0 && (blockIdx.x >> 1) == 0 && lane == 0 && trc_n < 2046u) {                               \
            a.trace[(role_) * 2048 + 1 + trc_n] = ((unsigned long long)(code_) << 56) | ((unsigned long long)clock64() & 0x00FFFFFFFFFFFFFFull); \
            a.trace[(role_) * 2048] = ++trc_n;                                                                          \
        }                                                                                                               \
    } while (0)
#define T2_DBG(a_, bit_) (((a_).debug & (bit_)) != 0)     // tc_debug: 2 = no global stores, 32 = no residual loads, 64 = no TMA store / bulk-group instructions at all (timing experiments)
#else
#define T2_TRACE(role_, code_) do { } while (0)
#define T2_DBG(a_, bit_) false
#endif

constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;       // shared::cluster address of the same offset in the pair's EVEN (leader) CTA
constexpr int T2_TW = 8, T2_TH = 16;
constexpr int T2_MAX_SLOTS = 8;
constexpr int T2_TMEM_COLS = 512;
constexpr int B2_AFULL = 0, B2_TFULL = 8, B2_TEMPTY = 16, B2_BRES = 24, B2_TMEMPTR = 26, B2_RFULL = 28, B2_PARAMS = 76;   // uint64 slots
constexpr int T2_STAGE_BYTES = 1024;              // one epilogue item: 32 pixels x 16 channels bf16, staged for the TMA store
constexpr int T2_NBUF = 3;                        // staging buffers per epilogue warp (a residual tile is loaded one item ahead)

struct Tc2Args {
    int B, H, W, Cin, Cout;
    int ksize, pad;
    int n_tile;                               // 2 * Cout
    int tiles_x, tiles_y;
    long long n_tiles;                        // M tiles of the layer
    float inv_tx, inv_ty;
    int slots;                                // A ring depth == accumulator ring depth (resident weights)
    int kchunks, nn_log2, a_stages, b_stages; // streamed weights: K chunks of 64 channels, log2(n tiles), ring depths
    int halo_w;
    uint32_t a_tx_bytes, a_bytes;             // halo tile bytes, rounded to 1 KB
    uint32_t b_half_bytes;                    // one tap's half weight tile: (n_tile / 2) x cin_blk bf16
    uint32_t b_region_off;
    uint32_t stage_off;                       // 16 warps x T2_NBUF x 1 KB staging buffers of the epilogue's TMA stores
    int elu, pdl, tma_out;
    int reverse;                              // walk the units from the last to the first (read_conv_plan_set_tile_order)
    const float *bias_f, *bias_m, *scale, *shift;
    const __nv_bfloat16 *residual;
    __nv_bfloat16 *out;
    unsigned long long *trace;                // READ_DIAG builds only
    int debug;                                // READ_DIAG builds only
};

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA tile loads whose completion is signalled on the LEADER's mbarrier (the data lands in the issuing CTA's shared memory)
__device__ __forceinline__ void tma2_load_4d(const CUtensorMap *tm, uint32_t bar_leader, uint32_t dst, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar_leader), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap *tm, uint32_t bar_leader, uint32_t dst, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar_leader), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// D[tmem, 256 rows over the CTA pair] (+)= A * B, one issuing thread of the leader CTA
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                           uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, {%7, %7, %7, %7, %7, %7, %7, %7}, p;\n\t"
        "}" ::"r"(d_tmem),
        "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// mbarrier at the same offset in BOTH CTAs arrives once all tcgen05.mma issued so far by this thread have completed
__device__ __forceinline__ void umma2_commit_multicast(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((unsigned short)3)
                 : "memory");
}
__device__ __forceinline__ uint32_t cvt2_bf16x2(float lo, float hi)
{
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

struct Tile2 {
    int tx, ty, b;
};
__device__ __forceinline__ Tile2 decode2(int t, const Tc2Args &a)
{
    Tile2 c;
    const int q = fdiv_small(t, a.inv_tx);
    c.tx = t - q * a.tiles_x;
    c.b = fdiv_small(q, a.inv_ty);
    c.ty = q - c.b * a.tiles_y;
    return c;
}

// KS = filter size, KKN = cin_blk / 16, EPI: 0 = runtime flags, 1 = ELU / no residual, 2 = no activation + residual
template <int KS, int KKN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(640, 1)
gated_conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
                      const __grid_constant__ CUtensorMap tmR, const __grid_constant__ Tc2Args a)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (s_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem_al = smem_raw + (smem_base - s_u32(smem_raw));
    constexpr int ntaps = KS * KS;
    const uint32_t b_region = smem_base + a.b_region_off;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_al + a.stage_off + 16u * T2_NBUF * T2_STAGE_BYTES);
    const uint32_t bar0 = s_u32(bars);
    const uint32_t afull0 = bar0 + 8 * B2_AFULL, tfull0 = bar0 + 8 * B2_TFULL, tempty0 = bar0 + 8 * B2_TEMPTY, bres = bar0 + 8 * B2_BRES;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + B2_TMEMPTR);
    float *s_par = reinterpret_cast<float *>(bars + B2_PARAMS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if (a.pdl) pdl_launch_dependents();
#ifdef READ_DIAG
    unsigned trc_n = 0;
#endif

    for (int i = threadIdx.x; i < a.Cout; i += 640)        // {bias_f, bias_m / 2, bn_scale / 2, bn_shift}: gate_folded
        reinterpret_cast<float4 *>(s_par)[i] = make_float4(a.bias_f[i], 0.5f * a.bias_m[i], 0.5f * a.scale[i], a.shift[i]);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
        if (a.residual != nullptr) tma_prefetch_desc(&tmR);
    }
    const uint32_t items_per_tile = (uint32_t)(a.n_tile >> 3);              // 4 quadrants x nch16 chunks
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < T2_MAX_SLOTS; ++s) {
            mbar_init(afull0 + 8 * s, 1);
            mbar_init(tfull0 + 8 * s, 1);
            mbar_init(tempty0 + 8 * s, 2u * items_per_tile);               // both CTAs' epilogue items
        }
        mbar_init(bres, 1);
        for (int i = 0; i < 16 * T2_NBUF; ++i) mbar_init(bar0 + 8 * (B2_RFULL + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc2(s_u32(tmem_ptr_smem), T2_TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                       // barriers of both CTAs initialised, TMEM allocated in both
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const uint32_t n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
    const long long n_units = (a.n_tiles + 1) >> 1;
    const uint32_t my_units = (uint32_t)((n_units - cluster_id + n_clusters - 1) / n_clusters);   // same in both CTAs of the pair
    const uint32_t slots = (uint32_t)a.slots;
    // unit i of this cluster -> this CTA's tile; a.reverse walks the image bottom-up (see read_conv_plan_set_tile_order)
    auto tile_of = [&](uint32_t i) -> long long {
        long long u = (long long)cluster_id + (long long)i * n_clusters;
        if (a.reverse) u = n_units - 1 - u;
        return 2ll * u + rank;
    };

    if (warp == 0) {
        // ===================== TMA producer (one per CTA) =====================
        if (elect_one()) {
            // resident HALF of the weights: rows [rank * N/2, (rank + 1) * N/2) of every tap; both CTAs signal the leader's barrier
            if (leader) mbar_arrive_expect_tx(bres, 2u * (uint32_t)ntaps * a.b_half_bytes);
            for (int i = 0; i < ntaps; ++i)
                tma2_load_2d(&tmB, bres & PEER_MASK, b_region + (uint32_t)i * a.b_half_bytes, 0, i * a.n_tile + (int)rank * (a.n_tile >> 1));
        }
        __syncwarp();
        if (a.pdl) pdl_wait();                // activations come from the previous kernel; the (static) weights above do not
        uint32_t s = 0, ph = 0;
        for (uint32_t i = 0; i < my_units; ++i) {
            const long long t = tile_of(i);
            const Tile2 tc = decode2((int)t, a);      // t >= n_tiles (odd tile count): b == B, the load is zero-filled
            mbar_wait(tfull0 + 8 * s, ph ^ 1u);       // stage free: the pair's MMAs of the tile that used it are complete
            T2_TRACE(rank * 8, 1);
            if (elect_one()) {
                if (leader) mbar_arrive_expect_tx(afull0 + 8 * s, 2u * a.a_tx_bytes);
                tma2_load_4d(&tmA, (afull0 + 8 * s) & PEER_MASK, smem_base + s * a.a_bytes, 0, tc.tx * T2_TW - a.pad, tc.ty * T2_TH - a.pad, tc.b);
            }
            __syncwarp();
            T2_TRACE(rank * 8, 2);
            if (++s == slots) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1 || warp == 3) {
        // ===================== MMA issuers (leader CTA only): warps 1 and 3 take alternate units =====================
        // Role timeline of the single-issuer version (profiles/r02_role_timelines.md): 574 + 370 cycles of barrier waits between
        // the last MMA of a unit and the first of the next, during which the tensor pipe drains its short queue and idles
        // (36 MMAs took 2085 cycles to issue, i.e. the issue is back-pressured by execution).  With two issuing threads the other
        // one has finished its waits and sits at its first MMA while this one is still issuing.  A unit's slot is i % slots, so
        // the issuers own disjoint slots (slots is even) and each commit covers exactly its own thread's MMAs.
        if (leader) {
            const uint32_t me = warp == 3 ? 1u : 0u;
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.n_tile >> 3) << 17) | ((256u >> 4) << 24);
            constexpr uint32_t row_bytes = KKN * 16u * 2u;
            constexpr uint32_t layout_type = (KKN == 4) ? 2u : 4u;         // SWIZZLE_128B : SWIZZLE_64B
            const uint32_t desc_hi = (uint32_t)(make_kmajor_desc(0, (uint32_t)a.halo_w * row_bytes, layout_type) >> 32);
            const uint32_t desc_hi_b = (uint32_t)(make_kmajor_desc(0, 8u * row_bytes, layout_type) >> 32);
            constexpr uint32_t lo_lbo = 1u << 16;
            constexpr uint32_t px16 = row_bytes >> 4;
            const uint32_t ky_step = (uint32_t)a.halo_w * px16;
            const uint32_t a16 = a.a_bytes >> 4, b16 = a.b_half_bytes >> 4;
            const uint32_t a_lo0 = ((smem_base & 0x3FFFFu) >> 4) | lo_lbo, b_lo0 = ((b_region & 0x3FFFFu) >> 4) | lo_lbo;
            mbar_wait(bres, 0);
            uint32_t s = 0, ph = 0;
            for (uint32_t i = 0; i < my_units; ++i) {
                const uint32_t s_i = s, ph_i = ph;
                if (++s == slots) { s = 0; ph ^= 1u; }
                if ((i & 1u) != me) continue;                           // the other issuer's unit
                mbar_wait(tempty0 + 8 * s_i, ph_i ^ 1u);
                T2_TRACE(2 + me, 3);
                mbar_wait(afull0 + 8 * s_i, ph_i);
                T2_TRACE(2 + me, 4);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + s_i * (uint32_t)a.n_tile;
                const uint32_t a_lo = a_lo0 + s_i * a16;
                if (elect_one()) {
#pragma unroll
                    for (int kx = 0; kx < KS; ++kx) {
#pragma unroll
                        for (int ky = 0; ky < KS; ++ky) {
                            const uint32_t bl = b_lo0 + (uint32_t)(ky * KS + kx) * b16;
                            const uint32_t al = a_lo + (uint32_t)ky * ky_step + (uint32_t)kx * px16;
#pragma unroll
                            for (int kk = 0; kk < KKN; ++kk)
                                umma2_bf16(d_tmem, al + 2u * kk, desc_hi, bl + 2u * kk, desc_hi_b, idesc, (kx | ky | kk) != 0 ? 1u : 0u);
                        }
                    }
                    umma2_commit_multicast(tfull0 + 8 * s_i);
                }
                __syncwarp();
                T2_TRACE(2 + me, 5);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: lean item loop (see conv_tc.cu) over this CTA's tile of every unit =====================
        const int q = warp & 3;
        const int sub = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const int py = r / T2_TW, px = r % T2_TW;
        const int half = a.n_tile >> 1;
        const float4 *par4 = reinterpret_cast<const float4 *>(s_par);
        const bool elu = EPI == 1 ? true : (EPI == 2 ? false : a.elu != 0);
        const bool has_res = EPI == 2 ? true : (EPI == 1 ? false : a.residual != nullptr);
        const int nch16 = half >> 4;
        const int lg = nch16 >> 1;
        const int chunk = sub & (nch16 - 1);
        const uint32_t item_step = 4u >> lg;
        if (a.pdl) pdl_wait();
#ifdef READ_DIAG
        const int trole = (warp == 4 || warp == 5) ? (int)rank * 8 + warp : -1;
#endif
        uint32_t acc = ((uint32_t)sub >> lg) % slots, acc_ph = (((uint32_t)sub >> lg) / slots) & 1u;
        // Output path.  Round-2 timing experiments (scripts/ab_pair_dbg.py, profiles/r02_conv_experiments.md): with the epilogue's
        // global stores switched off a C=64 layer ran in 50 instead of 64 us - a lane's two 16-byte stores at a 128-byte lane stride
        // cost 32 LSU wavefronts per instruction.  The item (32 pixels x 16 channels = 1 KB) is therefore staged in a per-warp
        // shared-memory buffer (conflict-free with the 32-byte TMA swizzle) and written by ONE TMA store per item; a residual tile
        // is TMA-loaded into the same buffer one item ahead and updated in place.  No cross-warp synchronisation is involved: each
        // warp's lane 0 owns its bulk groups.  Ragged edges and the phantom tile of an odd tile count are clipped by the TMA unit.
        const uint32_t sbuf0 = smem_base + a.stage_off + (uint32_t)(warp - 4) * (T2_NBUF * T2_STAGE_BYTES);
        const uint32_t rfull0 = bar0 + 8 * (B2_RFULL + (warp - 4) * T2_NBUF);
        const uint32_t lane_off = (uint32_t)lane * 32u, sw = (((uint32_t)lane >> 2) & 1u) * 16u;     // SWIZZLE_32B: bit 4 ^= bit 7
        const int co = chunk * 16;
        uint32_t k = 0, kph = 0;                                                                      // staging buffer of this item
        const bool tma_out = a.tma_out != 0;
        if (tma_out && has_res && lane == 0 && ((uint32_t)sub >> lg) < my_units) {
            const long long t0 = tile_of((uint32_t)sub >> lg);
            const Tile2 t0c = decode2((int)t0, a);
            mbar_arrive_expect_tx(rfull0, T2_STAGE_BYTES);
            tma_load_4d(&tmR, rfull0, sbuf0, co, t0c.tx * T2_TW, t0c.ty * T2_TH + q * 4, t0c.b);
        }
        for (uint32_t it = (uint32_t)sub >> lg; it < my_units; it += item_step) {
            const long long t = tile_of(it);
            const Tile2 tc = decode2((int)t, a);
            const int b = tc.b;
            const int x = tc.tx * T2_TW + px, y = tc.ty * T2_TH + py;
            const bool inside = (t < a.n_tiles) && (x < a.W) && (y < a.H);
            const uint32_t trow = tmem_base + acc * (uint32_t)a.n_tile + ((uint32_t)(q * 32) << 16);
            const int o = ((b * a.H + y) * a.W + x) * a.Cout + co;
            uint4 rs0 = make_uint4(0, 0, 0, 0), rs1 = rs0;
            const uint32_t kn = k + 1 == T2_NBUF ? 0u : k + 1;
            if (tma_out) {
                if (lane == 0) {
                    if (!T2_DBG(a, 64)) bulk_wait_group_read<1>();          // only the previous item's store may still be reading: buffers k and kn are free
                    if (has_res && it + item_step < my_units) {
                        const long long tn = tile_of(it + item_step);
                        const Tile2 tnc = decode2((int)tn, a);
                        mbar_arrive_expect_tx(rfull0 + 8 * kn, T2_STAGE_BYTES);
                        tma_load_4d(&tmR, rfull0 + 8 * kn, sbuf0 + kn * T2_STAGE_BYTES, co, tnc.tx * T2_TW, tnc.ty * T2_TH + q * 4, tnc.b);
                    }
                }
            } else if (inside && has_res && !T2_DBG(a, 32)) {
                rs0 = __ldg(reinterpret_cast<const uint4 *>(a.residual + o));
                rs1 = __ldg(reinterpret_cast<const uint4 *>(a.residual + o) + 1);
            }
            T2_TRACE(trole, 9);
            mbar_wait(tfull0 + 8 * acc, acc_ph);
            T2_TRACE(trole, 6);
            tcgen05_fence_after();
            uint32_t f16[16], m16[16];
            tmem_ld16(trow + (uint32_t)(chunk * 16), f16);
            tmem_ld16(trow + (uint32_t)(half + chunk * 16), m16);
            tmem_ld_wait();
            T2_TRACE(trole, 7);
            // the accumulator is in registers: hand the TMEM slot back to the leader's issuer before the math
            tcgen05_fence_before();
            __syncwarp();                               // (also orders lane 0's wait_group.read before the other lanes' staging writes)
            if (lane == 0) mbar_arrive_cluster((tempty0 + 8 * acc) & PEER_MASK);
            float yv[16];
            if (elu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) yv[j] = gate_folded<true>(__uint_as_float(f16[j]), __uint_as_float(m16[j]), par4[co + j]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) yv[j] = gate_folded<false>(__uint_as_float(f16[j]), __uint_as_float(m16[j]), par4[co + j]);
            }
            T2_TRACE(trole, 10);
            if (tma_out) {
                const uint32_t sb = sbuf0 + k * T2_STAGE_BYTES + lane_off;
                if (has_res) {
                    mbar_wait(rfull0 + 8 * k, kph);
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rs0.x), "=r"(rs0.y), "=r"(rs0.z), "=r"(rs0.w) : "r"(sb + sw));
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rs1.x), "=r"(rs1.y), "=r"(rs1.z), "=r"(rs1.w) : "r"(sb + (sw ^ 16u)));
                    const uint32_t rr[8] = {rs0.x, rs0.y, rs0.z, rs0.w, rs1.x, rs1.y, rs1.z, rs1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        yv[2 * j] += __uint_as_float(rr[j] << 16);
                        yv[2 * j + 1] += __uint_as_float(rr[j] & 0xFFFF0000u);
                    }
                }
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) pk[j] = cvt2_bf16x2(yv[2 * j], yv[2 * j + 1]);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + sw), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + (sw ^ 16u)), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                T2_TRACE(trole, 11);
                fence_proxy_async_smem();
                __syncwarp();
                T2_TRACE(trole, 12);
                if (lane == 0 && !T2_DBG(a, 64)) {      // (64: timing experiment without any bulk-group bookkeeping)
                    if (!T2_DBG(a, 2)) tma_store_4d(&tmO, sbuf0 + k * T2_STAGE_BYTES, co, tc.tx * T2_TW, tc.ty * T2_TH + q * 4, b);
                    bulk_commit_group();
                }
                k = kn;
                if (k == 0) kph ^= 1u;
            } else if (inside && !T2_DBG(a, 2)) {
                if (has_res) {
                    const uint32_t rr[8] = {rs0.x, rs0.y, rs0.z, rs0.w, rs1.x, rs1.y, rs1.z, rs1.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        yv[2 * j] += __uint_as_float(rr[j] << 16);
                        yv[2 * j + 1] += __uint_as_float(rr[j] & 0xFFFF0000u);
                    }
                }
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) pk[j] = cvt2_bf16x2(yv[2 * j], yv[2 * j + 1]);
                uint4 *op = reinterpret_cast<uint4 *>(a.out + o);
                op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                op[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
            T2_TRACE(trole, 8);
            acc += item_step;
            while (acc >= slots) { acc -= slots; acc_ph ^= 1u; }
        }
        if (tma_out && lane == 0) bulk_wait_group<0>();          // the staging buffers are read, the stores performed, before the CTA retires
    }

    // nobody leaves while the peer may still signal this CTA's barriers or its TMEM is in use
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc2(tmem_base, T2_TMEM_COLS);
    }
}

// ------------------------------------------------------------------ streamed-weight variant (Cin, Cout = 128 / 256)
// The wide ResBlock layers do not fit their weights in shared memory: the single-CTA kernel streams one [n_tile x 64] weight tile
// (32 KB) per tap and K chunk for every 128-pixel tile, ~590 KB per tile, 600 MB per layer - at the layer's 66 us that is 9 TB/s out
// of L2, i.e. the layer is L2-bandwidth bound, not tensor bound (72 MMAs x 128 cycles per tile would be 38 us).  As a CTA pair
// (M = 256) each CTA streams only ITS half of the weight rows (CTA 0: conv_f rows, CTA 1: conv_m rows of the n tile), halving the
// L2 -> SM traffic per pixel.  Work unit = (tile pair, n tile); rings: A = one halo tile per (unit, K chunk), B = one half weight
// tile per (unit, K chunk, tap); accumulators: 512 / n_tile TMEM slots.  Unit width n_tile = 256 columns.  (Cout = 256 at C3: its
// 1/8-size image has only 128 tile pairs -> 256 units of ~18 us over 74 clusters = 3.46 rounds of work in 4.  128-column units
// ("tc_wide_ntile" 128: 512 half-size units, 7 rounds of 6.9) were measured SLOWER, 73 vs 66 us per layer: with four 64-cycle MMAs per
// weight stage the issuing thread's per-stage wait + commit is no longer hidden.)  The weights stay in conv_tc's packing ([tap][K chunk]
// [n tile of 256: 128 conv_f rows | 128 conv_m rows][64]): a unit's conv_f / conv_m rows are two row ranges of it.
//   afull / bfull (leader's): both CTAs' TMA loads complete_tx on the leader's barrier
//   aempty / bempty / tfull (each CTA's own): tcgen05.commit multicast
//   tempty (leader's): every epilogue warp of both CTAs arrives after its last TMEM load of the unit
constexpr int S2_AFULL = 0, S2_AEMPTY = 4, S2_BFULL = 8, S2_BEMPTY = 24, S2_TFULL = 40, S2_TEMPTY = 44, S2_TMEMPTR = 48, S2_PARAMS = 50;
constexpr int S2_MAX_A = 4, S2_MAX_B = 16;

template <int KS, int KKN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(640, 1)
gated_conv_tc2s_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ Tc2Args a)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (s_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem_al = smem_raw + (smem_base - s_u32(smem_raw));
    constexpr int ntaps = KS * KS;
    const uint32_t b_region = smem_base + a.b_region_off;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_al + a.b_region_off + (uint32_t)a.b_stages * (uint32_t)KS * a.b_half_bytes);
    const uint32_t bar0 = s_u32(bars);
    const uint32_t afull0 = bar0 + 8 * S2_AFULL, aempty0 = bar0 + 8 * S2_AEMPTY, bfull0 = bar0 + 8 * S2_BFULL, bempty0 = bar0 + 8 * S2_BEMPTY;
    const uint32_t tfull0 = bar0 + 8 * S2_TFULL, tempty0 = bar0 + 8 * S2_TEMPTY;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + S2_TMEMPTR);
    float *s_par = reinterpret_cast<float *>(bars + S2_PARAMS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if (a.pdl) pdl_launch_dependents();

    for (int i = threadIdx.x; i < a.Cout; i += 640)        // {bias_f, bias_m, bn_scale, bn_shift}: gate_fast, as the single-CTA wide epilogue
        reinterpret_cast<float4 *>(s_par)[i] = make_float4(a.bias_f[i], a.bias_m[i], a.scale[i], a.shift[i]);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < S2_MAX_A; ++s) {
            mbar_init(afull0 + 8 * s, 1);
            mbar_init(aempty0 + 8 * s, 1);
        }
        for (int s = 0; s < S2_MAX_B; ++s) {
            mbar_init(bfull0 + 8 * s, 1);
            mbar_init(bempty0 + 8 * s, 1);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(tfull0 + 8 * s, 1);
            mbar_init(tempty0 + 8 * s, 32);                  // 16 epilogue warps of each CTA
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc2(s_u32(tmem_ptr_smem), T2_TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const uint32_t n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
    const int nn_log2 = a.nn_log2;                           // log2(units per tile pair) = log2(2 * Cout / n_tile): 0, 1 or 2
    const uint32_t slots = (uint32_t)a.slots;                // accumulator ring: 512 / n_tile
    const long long n_units = ((a.n_tiles + 1) >> 1) << nn_log2;
    const uint32_t my_units = n_units > cluster_id ? (uint32_t)((n_units - cluster_id + n_clusters - 1) / n_clusters) : 0u;
    const int half = a.n_tile >> 1;
    const int n_total = 2 * a.Cout;
    const uint32_t a_stages = (uint32_t)a.a_stages, b_stages = (uint32_t)a.b_stages;
    const int kchunks = a.kchunks;
    // unit i of this cluster -> (tile of this CTA, n tile)
    auto unit_tile = [&](uint32_t i, int &nt) -> long long {
        long long u = (long long)cluster_id + (long long)i * n_clusters;
        if (a.reverse) u = n_units - 1 - u;
        nt = (int)(u & ((1 << nn_log2) - 1));
        return 2ll * (u >> nn_log2) + rank;
    };

    if (warp == 0) {
        // ===================== A producer: one halo tile per (unit, K chunk) =====================
        if (a.pdl) pdl_wait();
        uint32_t as = 0, aph = 0;
        for (uint32_t i = 0; i < my_units; ++i) {
            int nt;
            const long long t = unit_tile(i, nt);
            const Tile2 tc = decode2((int)t, a);
            for (int kc = 0; kc < kchunks; ++kc) {
                mbar_wait(aempty0 + 8 * as, aph ^ 1u);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(afull0 + 8 * as, 2u * a.a_tx_bytes);
                    tma2_load_4d(&tmA, (afull0 + 8 * as) & PEER_MASK, smem_base + as * a.a_bytes, kc * (KKN * 16), tc.tx * T2_TW - a.pad,
                                 tc.ty * T2_TH - a.pad, tc.b);
                }
                __syncwarp();
                if (++as == a_stages) { as = 0; aph ^= 1u; }
            }
        }
    } else if (warp == 2) {
        // ===================== B producer: this CTA's half of the weight rows, one tile per (unit, K chunk, tap) =====================
        uint32_t bs = 0, bph = 0;
        for (uint32_t i = 0; i < my_units; ++i) {
            int nt;
            (void)unit_tile(i, nt);
            // conv_tc packing: n tile T of 256 rows = 128 conv_f rows then 128 conv_m rows; this unit's channels start at nt * half
            const int row0 = ((nt * half) >> 7) * 256 + (int)rank * 128 + ((nt * half) & 127);
            for (int kc = 0; kc < kchunks; ++kc) {
                for (int ky = 0; ky < KS; ++ky) {                  // one stage = the KS taps of a filter row
                    mbar_wait(bempty0 + 8 * bs, bph ^ 1u);
                    if (elect_one()) {
                        if (leader) mbar_arrive_expect_tx(bfull0 + 8 * bs, 2u * (uint32_t)KS * a.b_half_bytes);
#pragma unroll
                        for (int kx = 0; kx < KS; ++kx)
                            tma2_load_2d(&tmB, (bfull0 + 8 * bs) & PEER_MASK, b_region + (bs * (uint32_t)KS + (uint32_t)kx) * a.b_half_bytes, 0,
                                         ((ky * KS + kx) * kchunks + kc) * n_total + row0);
                    }
                    __syncwarp();
                    if (++bs == b_stages) { bs = 0; bph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.n_tile >> 3) << 17) | ((256u >> 4) << 24);
            constexpr uint32_t row_bytes = KKN * 16u * 2u;
            constexpr uint32_t layout_type = (KKN == 4) ? 2u : 4u;
            const uint32_t desc_hi = (uint32_t)(make_kmajor_desc(0, (uint32_t)a.halo_w * row_bytes, layout_type) >> 32);
            const uint32_t desc_hi_b = (uint32_t)(make_kmajor_desc(0, 8u * row_bytes, layout_type) >> 32);
            constexpr uint32_t lo_lbo = 1u << 16;
            constexpr uint32_t px16 = row_bytes >> 4;
            const uint32_t ky_step = (uint32_t)a.halo_w * px16;
            const uint32_t a16 = a.a_bytes >> 4, b16 = a.b_half_bytes >> 4;
            const uint32_t a_lo0 = ((smem_base & 0x3FFFFu) >> 4) | lo_lbo, b_lo0 = ((b_region & 0x3FFFFu) >> 4) | lo_lbo;
            uint32_t as = 0, aph = 0, bs = 0, bph = 0, acc = 0, acc_ph = 0;
            for (uint32_t i = 0; i < my_units; ++i) {
                mbar_wait(tempty0 + 8 * acc, acc_ph ^ 1u);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)a.n_tile;
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(afull0 + 8 * as, aph);
                    tcgen05_fence_after();
                    const uint32_t a_lo = a_lo0 + as * a16;
#pragma unroll
                    for (int ky = 0; ky < KS; ++ky) {
                        // one weight stage = a filter row (KS taps, KS * KKN MMAs): with a stage per TAP the issuing thread's wait +
                        // commit per four MMAs showed in the tensor pipe's duty cycle (80 %, profiles/r02_ncu_full.md)
                        mbar_wait(bfull0 + 8 * bs, bph);
                        tcgen05_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int kx = 0; kx < KS; ++kx) {
                                const uint32_t bl = b_lo0 + (bs * (uint32_t)KS + (uint32_t)kx) * b16;
                                const uint32_t al = a_lo + (uint32_t)ky * ky_step + (uint32_t)kx * px16;
#pragma unroll
                                for (int kk = 0; kk < KKN; ++kk)
                                    umma2_bf16(d_tmem, al + 2u * kk, desc_hi, bl + 2u * kk, desc_hi_b, idesc, (kc | ky | kx | kk) != 0 ? 1u : 0u);
                            }
                            umma2_commit_multicast(bempty0 + 8 * bs);
                        }
                        __syncwarp();
                        if (++bs == b_stages) { bs = 0; bph ^= 1u; }
                    }
                    if (elect_one()) {
                        umma2_commit_multicast(aempty0 + 8 * as);
                        if (kc == kchunks - 1) umma2_commit_multicast(tfull0 + 8 * acc);
                    }
                    __syncwarp();
                    if (++as == a_stages) { as = 0; aph ^= 1u; }
                }
                if (++acc == slots) { acc = 0; acc_ph ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: 16 warps, warp (q, sub) owns the 16-channel chunks sub, sub + 4, .. of its 32 pixels =====================
        const int q = warp & 3;
        const int sub = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const int py = r / T2_TW, px = r % T2_TW;
        const float4 *par4 = reinterpret_cast<const float4 *>(s_par);
        const int nch16 = half >> 4;
        if (a.pdl) pdl_wait();
        uint32_t acc = 0, acc_ph = 0;
        for (uint32_t i = 0; i < my_units; ++i) {
            int nt;
            const long long t = unit_tile(i, nt);
            const Tile2 tc = decode2((int)t, a);
            const int x = tc.tx * T2_TW + px, y = tc.ty * T2_TH + py;
            const bool inside = (t < a.n_tiles) && (x < a.W) && (y < a.H);
            const uint32_t trow = tmem_base + acc * (uint32_t)a.n_tile + ((uint32_t)(q * 32) << 16);
            const int o0 = ((tc.b * a.H + y) * a.W + x) * a.Cout + nt * half;
            mbar_wait(tfull0 + 8 * acc, acc_ph);
            tcgen05_fence_after();
            for (int c = sub; c < nch16; c += 4) {
                const int o = o0 + c * 16;
                uint4 rs0 = make_uint4(0, 0, 0, 0), rs1 = rs0;
                if (inside && a.residual != nullptr) {
                    rs0 = __ldg(reinterpret_cast<const uint4 *>(a.residual + o));
                    rs1 = __ldg(reinterpret_cast<const uint4 *>(a.residual + o) + 1);
                }
                uint32_t f16[16], m16[16];
                tmem_ld16(trow + (uint32_t)(c * 16), f16);
                tmem_ld16(trow + (uint32_t)(half + c * 16), m16);
                tmem_ld_wait();
                if (c + 4 >= nch16) {          // last chunk of this warp: the accumulator slot goes back to the issuer before the math
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster((tempty0 + 8 * acc) & PEER_MASK);
                }
                const int co = nt * half + c * 16;
                float yv[16];
                if (a.elu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float4 pp = par4[co + j];
                        yv[j] = gate_fast<true>(__uint_as_float(f16[j]) + pp.x, __uint_as_float(m16[j]) + pp.y, pp.z, pp.w);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float4 pp = par4[co + j];
                        yv[j] = gate_fast<false>(__uint_as_float(f16[j]) + pp.x, __uint_as_float(m16[j]) + pp.y, pp.z, pp.w);
                    }
                }
                if (inside) {
                    if (a.residual != nullptr) {
                        const uint32_t rr[8] = {rs0.x, rs0.y, rs0.z, rs0.w, rs1.x, rs1.y, rs1.z, rs1.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            yv[2 * j] += __uint_as_float(rr[j] << 16);
                            yv[2 * j + 1] += __uint_as_float(rr[j] & 0xFFFF0000u);
                        }
                    }
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) pk[j] = cvt2_bf16x2(yv[2 * j], yv[2 * j + 1]);
                    uint4 *op = reinterpret_cast<uint4 *>(a.out + o);
                    op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    op[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
            if (++acc == slots) { acc = 0; acc_ph ^= 1u; }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc2(tmem_base, T2_TMEM_COLS);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();           // conv_tc.cu

struct Tc2Plan {
    CUtensorMap tmA, tmB, tmO, tmR;
    Tc2Args args;
    size_t smem_bytes;
    int kkn, epi, wide;
    int reverse;
};

int g_tc_wide_ntile = 256;       // read_set_option "tc_wide_ntile": unit width of the streamed pair kernel for Cout = 256 (128: measured slower)

bool tc2_supported(const read_conv_desc &d)
{
    if (d.act_dtype != READ_ACT_BF16 || d.mul != nullptr || d.n_src != 1 || d.src[0].mode != READ_SRC_IDENTITY) return false;
    if (d.stride != 1 || d.k != 3 || d.pad != 1 || d.Hin != d.Hout || d.Win != d.Wout) return false;
    const bool wide = (d.Cin == 128 || d.Cin == 256) && (d.Cout == 128 || d.Cout == 256);     // streamed weights, 256-column n tiles
    if (!wide) {
        if (!(d.Cin == 32 || d.Cin == 64)) return false;                       // one K chunk, resident weights
        if (!(d.Cout == 16 || d.Cout == 32 || d.Cout == 64)) return false;
    }
    if (d.out_mode != READ_OUT_NHWC || d.out2 != nullptr || d.addin != nullptr) return false;
    if ((long long)d.B * d.Hout * d.Wout * d.Cout >= (1ll << 31)) return false;
    return true;
}

int tc2_plan_create(const read_conv_desc &d, Tc2Plan **out)
{
    if (!tc2_supported(d)) {
        set_error("tcgen05 pair conv: unsupported layer");
        return READ_ERR_UNSUPPORTED;
    }
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) {
        set_error("tcgen05 pair conv: cuTensorMapEncodeTiled not available from the driver");
        return READ_ERR_CUDA;
    }
    Tc2Plan *p = new (std::nothrow) Tc2Plan{};
    RB_CHECK_ARG(p != nullptr, "tcgen05 pair conv: out of host memory");
    const bool wide = d.Cin > 64;
    const int cin_blk = wide ? 64 : d.Cin, n_tile = wide ? (g_tc_wide_ntile == 128 && d.Cout == 256 ? 128 : 256) : 2 * d.Cout;
    const int halo_rows = T2_TH + d.k - 1, halo_w = T2_TW + d.k - 1;
    const CUtensorMapSwizzle sw = cin_blk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    {
        const read_src &sv = d.src[0];
        cuuint64_t dims[4] = {(cuuint64_t)sv.C, (cuuint64_t)sv.W, (cuuint64_t)sv.H, (cuuint64_t)d.B};
        cuuint64_t strides[3] = {(cuuint64_t)sv.C * 2, (cuuint64_t)sv.W * sv.C * 2, (cuuint64_t)sv.H * sv.W * sv.C * 2};
        cuuint32_t box[4] = {(cuuint32_t)cin_blk, (cuuint32_t)halo_w, (cuuint32_t)halo_rows, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&p->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(sv.ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tcgen05 pair conv: cuTensorMapEncodeTiled(activations) failed with %d", (int)r);
            delete p;
            return READ_ERR_CUDA;
        }
    }
    {   // weights packed as for conv_tc ([tap][n][cin_blk], conv_f rows then conv_m rows): box = HALF the N rows of one tap
        const cuuint64_t rows = (cuuint64_t)d.k * d.k * (d.Cin / cin_blk) * 2 * d.Cout;
        cuuint64_t dims[2] = {(cuuint64_t)cin_blk, rows};
        cuuint64_t strides[1] = {(cuuint64_t)cin_blk * 2};
        cuuint32_t box[2] = {(cuuint32_t)cin_blk, (cuuint32_t)(n_tile / 2)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&p->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(d.w_tc), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tcgen05 pair conv: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
            delete p;
            return READ_ERR_CUDA;
        }
    }
    if (!wide) {   // epilogue items: 16 channels x 8 x 4 pixels of the NHWC output (store) / of the residual tensor (load), 32-byte swizzle
        cuuint64_t dims[4] = {(cuuint64_t)d.Cout, (cuuint64_t)d.Wout, (cuuint64_t)d.Hout, (cuuint64_t)d.B};
        cuuint64_t strides[3] = {(cuuint64_t)d.Cout * 2, (cuuint64_t)d.Wout * d.Cout * 2, (cuuint64_t)d.Hout * d.Wout * d.Cout * 2};
        cuuint32_t box[4] = {16, (cuuint32_t)T2_TW, 4, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&p->tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
            r = enc(&p->tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d.residual ? const_cast<void *>(d.residual) : d.out, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("tcgen05 pair conv: cuTensorMapEncodeTiled(output items) failed with %d", (int)r);
            delete p;
            return READ_ERR_CUDA;
        }
    }
    Tc2Args &a = p->args;
    a.B = d.B; a.H = d.Hout; a.W = d.Wout; a.Cin = d.Cin; a.Cout = d.Cout;
    a.ksize = d.k; a.pad = d.pad;
    a.n_tile = n_tile;
    a.tiles_x = (d.Wout + T2_TW - 1) / T2_TW;
    a.tiles_y = (d.Hout + T2_TH - 1) / T2_TH;
    a.n_tiles = (long long)a.tiles_x * a.tiles_y * d.B;
    a.inv_tx = 1.0f / (float)a.tiles_x;
    a.inv_ty = 1.0f / (float)a.tiles_y;
    a.halo_w = halo_w;
    a.a_tx_bytes = (uint32_t)halo_rows * halo_w * cin_blk * 2u;
    a.a_bytes = (a.a_tx_bytes + 1023u) & ~1023u;
    a.b_half_bytes = (uint32_t)(n_tile / 2) * cin_blk * 2u;
    a.kchunks = d.Cin / cin_blk;
    a.nn_log2 = 0;
    while ((n_tile << a.nn_log2) < 2 * d.Cout) ++a.nn_log2;
    a.elu = d.elu;
    a.bias_f = d.bias_f; a.bias_m = d.bias_m; a.scale = d.bn_scale; a.shift = d.bn_shift;
    a.residual = static_cast<const __nv_bfloat16 *>(d.residual);
    a.out = static_cast<__nv_bfloat16 *>(d.out);
    p->kkn = cin_blk / 16;
    p->wide = wide ? 1 : 0;
    if (wide) {
        a.slots = T2_TMEM_COLS / n_tile;
        a.a_stages = 3;
        const size_t fixed = 1024 + 8 * S2_PARAMS + 16 * (size_t)d.Cout + 64;
        size_t left = 227 * 1024 - fixed - (size_t)a.a_stages * a.a_bytes;
        a.b_stages = (int)(left / ((size_t)d.k * a.b_half_bytes));          // a stage holds the k taps of one filter row
        if (a.b_stages > S2_MAX_B) a.b_stages = S2_MAX_B;
        if (a.b_stages < 2) {
            set_error("tcgen05 pair conv: layer does not fit shared memory");
            delete p;
            return READ_ERR_UNSUPPORTED;
        }
        a.b_region_off = (uint32_t)a.a_stages * a.a_bytes;
        a.stage_off = 0;
        p->smem_bytes = fixed + (size_t)a.a_stages * a.a_bytes + (size_t)a.b_stages * d.k * a.b_half_bytes;
        p->epi = 0;
        *out = p;
        return READ_OK;
    }
    const int nacc = T2_TMEM_COLS / n_tile > T2_MAX_SLOTS ? T2_MAX_SLOTS : T2_TMEM_COLS / n_tile;
    a.slots = nacc;
    {   // the A ring (one halo tile per accumulator slot) shares 227 KB with the resident weights and the epilogue's staging buffers
        const size_t fixed = 1024 + (size_t)d.k * d.k * a.b_half_bytes + 16 * T2_NBUF * T2_STAGE_BYTES + 8 * B2_PARAMS + 16 * (size_t)d.Cout + 64;
        while (a.slots > 2 && fixed + (size_t)a.slots * a.a_bytes > 227 * 1024) --a.slots;
    }
    a.b_region_off = (uint32_t)a.slots * a.a_bytes;
    a.stage_off = a.b_region_off + (uint32_t)(d.k * d.k) * a.b_half_bytes;          // a multiple of 1 KB (b_half_bytes is)
    p->smem_bytes = 1024 + (size_t)a.stage_off + 16 * T2_NBUF * T2_STAGE_BYTES + 8 * B2_PARAMS + 16 * (size_t)d.Cout + 64;
    p->epi = (a.elu && !a.residual) ? 1 : ((!a.elu && a.residual) ? 2 : 0);
    if (p->smem_bytes > 227 * 1024) {
        set_error("tcgen05 pair conv: layer does not fit shared memory");
        delete p;
        return READ_ERR_UNSUPPORTED;
    }
    *out = p;
    return READ_OK;
}

extern int g_tc_pdl;
extern unsigned long long *g_tc_trace;
extern int g_tc_debug;
int g_tc_tma_store = 1;          // read_set_option "tc_tma_store": epilogue output through staged TMA stores (0: per-lane global stores)

int tc2_plan_launch(const Tc2Plan *p, cudaStream_t st, int max_ctas)
{
    Tc2Args a = p->args;
    a.pdl = g_tc_pdl ? 1 : 0;
    a.trace = g_tc_trace;
    a.debug = g_tc_debug;
    a.tma_out = g_tc_tma_store ? 1 : 0;
    a.reverse = p->reverse;
    if (a.n_tiles == 0) return READ_OK;
    long long grid = num_sms() & ~1;                  // whole pairs
    if (max_ctas > 1 && grid > (max_ctas & ~1)) grid = max_ctas & ~1;
    const long long units = ((a.n_tiles + 1) / 2) << (p->wide ? a.nn_log2 : 0);
    if (grid > 2 * units) grid = 2 * units;
    cudaLaunchAttribute lattr[1];            // the cluster shape (2,1,1) is a compile-time attribute of the kernel
    lattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    lattr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t lcfg{};
    lcfg.gridDim = dim3((unsigned)grid);
    lcfg.blockDim = dim3(640);
    lcfg.dynamicSmemBytes = p->smem_bytes;
    lcfg.stream = st;
    lcfg.attrs = lattr;
    lcfg.numAttrs = a.pdl ? 1 : 0;
    if (p->wide) {
        RB_CUDA(cudaFuncSetAttribute(gated_conv_tc2s_kernel<3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
        RB_CUDA(cudaLaunchKernelEx(&lcfg, gated_conv_tc2s_kernel<3, 4>, p->tmA, p->tmB, a));
        RB_LAUNCH_CHECK();
        return READ_OK;
    }
#define RB_TC2(KKN_, EPI_)                                                                                              \
    do {                                                                                                                \
        RB_CUDA(cudaFuncSetAttribute(gated_conv_tc2_kernel<3, KKN_, EPI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)p->smem_bytes));                                                              \
        RB_CUDA(cudaLaunchKernelEx(&lcfg, gated_conv_tc2_kernel<3, KKN_, EPI_>, p->tmA, p->tmB, p->tmO, p->tmR, a));                    \
    } while (0)
    if (p->kkn == 2) {
        if (p->epi == 1) RB_TC2(2, 1); else if (p->epi == 2) RB_TC2(2, 2); else RB_TC2(2, 0);
    } else {
        if (p->epi == 1) RB_TC2(4, 1); else if (p->epi == 2) RB_TC2(4, 2); else RB_TC2(4, 0);
    }
#undef RB_TC2
    RB_LAUNCH_CHECK();
    return READ_OK;
}

void tc2_plan_destroy(Tc2Plan *p) { delete p; }
void tc2_plan_set_reverse(Tc2Plan *p, int reverse) { p->reverse = reverse ? 1 : 0; }

}  // namespace rb
