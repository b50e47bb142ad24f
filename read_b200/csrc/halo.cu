// Strip-parallel refinement net: halo exchange between neighbouring GPUs over NVLink peer memory.
//
// The frame is cut into horizontal strips, one per GPU; every activation tensor of a rank is the crop [y0 - h, y1 + h) of the
// global tensor (h halo rows, 16 at full resolution, halved per level).  A 3x3 conv invalidates one halo row per side, so
// before the valid halo runs out the engine exchanges boundary rows with the strip above / below (read_b200/engine.py:
// StripEngine).  One kernel per exchange and rank, no NCCL on the path:
//   phase 1  push my top / bottom interior rows into the NEIGHBOUR's mailbox (plain 16-byte stores through the peer mapping of
//            the neighbour's cudaMalloc'ed mailbox, opened with cudaIpcOpenMemHandle), __threadfence_system, and the last CTA to
//            finish publishes the frame's epoch in the neighbour's flag word (st.release.sys);
//   phase 2  spin (ld.acquire.sys) until BOTH neighbours' epochs have arrived in MY flags, then copy my mailbox slots into the
//            halo rows of the local tensor.
// Every exchange has its own mailbox slots, flags and CTA counter.  A slot is reused one frame later; the neighbour cannot reach
// the same exchange of the next frame before it has received this rank's pushes of all LATER exchanges of the current frame,
// which are stream-ordered after this rank's phase 2 - so with >= 2 exchanges per frame (the engine asserts it) a slot is never
// overwritten while it is being drained.  All kernels are plain launches: the whole strip net replays as one CUDA graph.
#include "common.cuh"
#include <string.h>

namespace rb {

struct HaloArgs {
    // phase 1: sources in the local tensor, destinations in the neighbours' mailboxes (peer pointers), null = no such neighbour
    const uint4 *src_up, *src_dn;
    uint4 *peer_up_slot, *peer_dn_slot;
    unsigned *peer_up_flag, *peer_dn_flag;
    // phase 2: my mailbox slots -> halo rows of the local tensor
    const uint4 *slot_from_up, *slot_from_dn;
    const unsigned *flag_from_up, *flag_from_dn;
    uint4 *dst_top, *dst_bot;
    long long n16;                  // 16-byte units per direction
    const unsigned *epoch;          // device-resident frame counter (read_epoch_bump)
    unsigned *cta_counter;          // one word per exchange, zero between launches
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256) halo_exchange_kernel(const __grid_constant__ HaloArgs a)
{
    const unsigned epoch = *a.epoch;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    // ---- phase 1: push
    if (a.peer_up_slot != nullptr)
        for (long long i = i0; i < a.n16; i += stride) a.peer_up_slot[i] = a.src_up[i];
    if (a.peer_dn_slot != nullptr)
        for (long long i = i0; i < a.n16; i += stride) a.peer_dn_slot[i] = a.src_dn[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(a.cta_counter, 1u);
        if (prev == gridDim.x - 1) {                 // every CTA's stores are fenced: publish
            *a.cta_counter = 0u;
            __threadfence_system();
            if (a.peer_up_flag != nullptr) st_release_sys(a.peer_up_flag, epoch);
            if (a.peer_dn_flag != nullptr) st_release_sys(a.peer_dn_flag, epoch);
        }
        // ---- phase 2: wait for the neighbours (bounded: a protocol bug traps instead of hanging the GPU)
        unsigned long long spins = 0;
        while ((a.flag_from_up != nullptr && ld_acquire_sys(a.flag_from_up) < epoch) ||
               (a.flag_from_dn != nullptr && ld_acquire_sys(a.flag_from_dn) < epoch)) {
            if (++spins > (1ull << 31)) __trap();
        }
    }
    __syncthreads();
    if (a.slot_from_up != nullptr)
        for (long long i = i0; i < a.n16; i += stride) a.dst_top[i] = __ldcv(a.slot_from_up + i);
    if (a.slot_from_dn != nullptr)
        for (long long i = i0; i < a.n16; i += stride) a.dst_bot[i] = __ldcv(a.slot_from_dn + i);
}

__global__ void epoch_bump_kernel(unsigned *epoch) { *epoch += 1u; }

}  // namespace rb

using namespace rb;

extern "C" {

int read_ipc_alloc(int64_t bytes, void **dev_ptr, unsigned char *handle64)
{
    RB_CHECK_ARG(bytes > 0 && dev_ptr && handle64, "ipc_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void *p = nullptr;
    RB_CUDA(cudaMalloc(&p, (size_t)bytes));
    RB_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        cudaFree(p);
        return READ_ERR_CUDA;
    }
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return READ_OK;
}

int read_ipc_open(const unsigned char *handle64, void **peer_ptr)
{
    RB_CHECK_ARG(handle64 && peer_ptr, "ipc_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    RB_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return READ_OK;
}

int read_ipc_close(void *peer_ptr)
{
    if (peer_ptr) RB_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return READ_OK;
}

int read_ipc_free(void *dev_ptr)
{
    if (dev_ptr) RB_CUDA(cudaFree(dev_ptr));
    return READ_OK;
}

int read_epoch_bump(uint32_t *epoch, void *stream)
{
    RB_CHECK_ARG(epoch != nullptr, "epoch_bump: null pointer");
    epoch_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(epoch);
    RB_LAUNCH_CHECK();
    return READ_OK;
}

int read_halo_exchange(const read_halo_desc *d, void *stream)
{
    RB_CHECK_ARG(d != nullptr && d->bytes > 0 && d->bytes % 16 == 0, "halo_exchange: bytes must be a positive multiple of 16");
    RB_CHECK_ARG(d->epoch && d->cta_counter, "halo_exchange: null epoch / counter");
    RB_CHECK_ARG((d->peer_up_slot == nullptr) == (d->src_up == nullptr) && (d->peer_dn_slot == nullptr) == (d->src_dn == nullptr),
                 "halo_exchange: a push needs source and destination");
    RB_CHECK_ARG((d->slot_from_up == nullptr) == (d->dst_top == nullptr) && (d->slot_from_dn == nullptr) == (d->dst_bot == nullptr),
                 "halo_exchange: a receive needs slot and destination");
    const void *ptrs[] = {d->src_up, d->src_dn, d->peer_up_slot, d->peer_dn_slot, d->slot_from_up, d->slot_from_dn, d->dst_top, d->dst_bot};
    for (const void *p : ptrs) RB_CHECK_ARG((reinterpret_cast<uintptr_t>(p) & 15) == 0, "halo_exchange: pointers must be 16-byte aligned");
    HaloArgs a{};
    a.src_up = (const uint4 *)d->src_up; a.src_dn = (const uint4 *)d->src_dn;
    a.peer_up_slot = (uint4 *)d->peer_up_slot; a.peer_dn_slot = (uint4 *)d->peer_dn_slot;
    a.peer_up_flag = (unsigned *)d->peer_up_flag; a.peer_dn_flag = (unsigned *)d->peer_dn_flag;
    a.slot_from_up = (const uint4 *)d->slot_from_up; a.slot_from_dn = (const uint4 *)d->slot_from_dn;
    a.flag_from_up = (const unsigned *)d->flag_from_up; a.flag_from_dn = (const unsigned *)d->flag_from_dn;
    a.dst_top = (uint4 *)d->dst_top; a.dst_bot = (uint4 *)d->dst_bot;
    a.n16 = d->bytes / 16;
    a.epoch = (const unsigned *)d->epoch;
    a.cta_counter = (unsigned *)d->cta_counter;
    long long ctas = (a.n16 + 255) / 256;
    if (ctas > 64) ctas = 64;                       // all CTAs spin in phase 2: keep them co-resident
    if (ctas < 1) ctas = 1;
    halo_exchange_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(a);
    RB_LAUNCH_CHECK();
    return READ_OK;
}

}  // extern "C"
