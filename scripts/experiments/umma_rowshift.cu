// Experiment: how does tcgen05.mma address the rows of a swizzled K-major A operand when the descriptor's start address
// is NOT aligned to the swizzle pattern (start inside the 8-row atom) and the stride between 8-row groups (SBO) is not a
// multiple of the pattern size?  Needed to load a conv halo tile ONCE and take the kx shifts through the descriptor.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I read_b200/csrc -o gpurun_out/umma_rowshift scripts/experiments/umma_rowshift.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace rb;

// pixel array: P pixels x ROWB bytes, element (p, c) at p*ROWB + ((c/8) ^ swz(p)) * 16 + (c%8)*2 with the ABSOLUTE-address
// swizzle TMA would apply (128B: chunk ^= p & 7; 64B: chunk ^= (p >> 1) & 3), value = (p % 32) * 8 + c/8
template <int ROWB>
__global__ void probe(int off_px, int group_px, int base_offset, int kind /*unused*/, float *out)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (s_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *S = smem_raw + (base - s_u32(smem_raw));
    constexpr int CH = ROWB / 2;               // channels per pixel (64 or 32)
    constexpr int P = 256;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    for (int i = threadIdx.x; i < P * CH; i += blockDim.x) {
        const int p = i / CH, c = i % CH;
        const int chunk = c / 8;
        const int sw = ROWB == 128 ? (p & 7) : ((p >> 1) & 3);
        const uint32_t o = (uint32_t)p * ROWB + (uint32_t)((chunk ^ sw) * 16) + (uint32_t)(c % 8) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(S + o) = __float2bfloat16((float)((p % 32) * 8 + chunk));
    }
    // B: 16 rows (n) x CH, canonical aligned swizzled layout at S + 64 KB: B[n][k] = 1 if k == (CH/16)*n ... select k = 4n (or 2n)
    uint8_t *Bs = S + 65536;
    for (int i = threadIdx.x; i < 16 * CH; i += blockDim.x) {
        const int n = i / CH, k = i % CH;
        const int chunk = k / 8;
        const int sw = ROWB == 128 ? (n & 7) : ((n >> 1) & 3);
        const uint32_t o = (uint32_t)n * ROWB + (uint32_t)((chunk ^ sw) * 16) + (uint32_t)(k % 8) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(Bs + o) = __float2bfloat16(k == (CH / 16) * n ? 1.f : 0.f);
    }
    if (threadIdx.x == 0) { mbar_init(s_u32(&bar), 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(s_u32(&tmem_ptr), 32);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tm = tmem_ptr;
    if (threadIdx.x == 0) {
        const uint32_t layout = ROWB == 128 ? 2u : 4u;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_addr = base + (uint32_t)off_px * ROWB;
        uint64_t adesc = make_kmajor_desc(a_addr, (uint32_t)group_px * ROWB, layout) | ((uint64_t)(base_offset & 7) << 49);
        uint64_t bdesc = make_kmajor_desc(base + 65536u, 8u * ROWB, layout);
        for (int kk = 0; kk < CH / 16; ++kk)
            umma_bf16(tm, adesc + 2u * kk, bdesc + 2u * kk, idesc, kk != 0);
        umma_commit(s_u32(&bar));
    }
    mbar_wait(s_u32(&bar), 0);
    tcgen05_fence_after();
    if (threadIdx.x < 128) {
        uint32_t r[16];
        tmem_ld16(tm + ((uint32_t)(threadIdx.x & ~31) << 16), r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[threadIdx.x * 16 + j] = __uint_as_float(r[j]);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 32);
}

template <int ROWB> void run(int off, int group_px, int bo)
{
    float *d; cudaMalloc(&d, 128 * 16 * 4);
    cudaFuncSetAttribute(probe<ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    probe<ROWB><<<1, 128, 100 * 1024>>>(off, group_px, bo, 0, d);
    cudaError_t e = cudaDeviceSynchronize();
    float h[128 * 16];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    // expected under "linear rows + absolute-address swizzle": row r -> pixel off + (r/8)*group_px + r%8, chunk n*? /8
    constexpr int CH = ROWB / 2;
    int ok = 0, tot = 0;
    printf("ROWB=%d off=%d group_px=%d base_offset=%d err=%s\n", ROWB, off, group_px, bo, cudaGetErrorString(e));
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 16; ++n) {
            const int p = off + (r / 8) * group_px + (r % 8);
            const int c = (CH / 16) * n;
            const float want = (float)((p % 32) * 8 + c / 8);
            ++tot; ok += (h[r * 16 + n] == want);
        }
    printf("   match with linear-rows/absolute-swizzle model: %d / %d\n", ok, tot);
    for (int r = 0; r < 20; ++r) {
        printf("   r=%3d:", r);
        for (int n = 0; n < 16; n += 2) { const int v = (int)h[r * 16 + n]; printf(" (p%%32=%2d,ch=%d)", v / 8, v % 8); }
        printf("\n");
    }
    cudaFree(d);
}

int main()
{
    for (int off : {0, 1, 3, 8, 10}) for (int bo : {0, -1}) {
        run<128>(off, 8, bo < 0 ? (off & 7) : 0);
        if ((off & 7) == 0) break;
    }
    run<128>(0, 10, 0); run<128>(1, 10, 0); run<128>(1, 10, 1); run<128>(12, 10, 0); run<128>(12, 10, 4);
    run<64>(0, 8, 0); run<64>(1, 8, 0); run<64>(1, 8, 1); run<64>(2, 8, 0); run<64>(0, 10, 0); run<64>(1, 10, 0); run<64>(11, 10, 0);
    return 0;
}
