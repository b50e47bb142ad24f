// Descriptor gather / scatter-add for sm_100a.
// Replaces PointTexture.forward and its autograd backward (READ/models/texture.py:42-70):
//   feat[b,c,y,x] = texture_[0,c,(int64)idx[b,0,y,x]]      (empty pixel carries idx 0 -> point 0)
// Descriptors are read from a point-major [N,D] shadow so a pixel touches one 32-byte sector.
#include "common.cuh"

namespace rb {

__global__ void tex_to_point_major_kernel(const float *__restrict__ cn, int D, long long N, float *__restrict__ nd)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N;
         i += (long long)gridDim.x * blockDim.x) {
        if (D == 8) {
            float v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = __ldg(cn + (long long)c * N + i);
            float4 *o = reinterpret_cast<float4 *>(nd + i * 8);
            o[0] = make_float4(v[0], v[1], v[2], v[3]);
            o[1] = make_float4(v[4], v[5], v[6], v[7]);
        } else {
            for (int c = 0; c < D; ++c) nd[i * D + c] = __ldg(cn + (long long)c * N + i);
        }
    }
}

__global__ void tex_to_channel_major_kernel(const float *__restrict__ nd, int D, long long N, float *__restrict__ cn)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N;
         i += (long long)gridDim.x * blockDim.x) {
        for (int c = 0; c < D; ++c) cn[(long long)c * N + i] = nd[i * D + c];
    }
}

__device__ __forceinline__ float tex_act(float v, int act)
{
    if (act == READ_TEXACT_SIGMOID) return 1.f / (1.f + expf(-v));
    if (act == READ_TEXACT_TANH) return tanhf(v);
    return v;
}

// SRC: 0 = float index map, 1 = packed zbuf
template <int SRC, int LAYOUT>
__global__ void gather_kernel(const float *__restrict__ tex, int D, long long N, const void *__restrict__ src, int B,
                              int h, int w, int act, void *__restrict__ out)
{
    const long long hw = (long long)h * w;
    const long long total = (long long)B * hw;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        long long id;
        if (SRC == 0) {
            id = (long long)static_cast<const float *>(src)[p];           // texture.py:52 .long()
        } else {
            const unsigned long long k = static_cast<const unsigned long long *>(src)[p];
            id = (k == ZBUF_EMPTY) ? 0ll : (long long)(k & 0xFFFFFFFFull);
        }
        // the reference does not bounds-check (index_select would raise); clamp to stay memory-safe
        if (id < 0) id = 0;
        if (id >= N) id = N - 1;
        const long long b = p / hw, q = p - b * hw;
        if (D == 8) {
            const float4 *t = reinterpret_cast<const float4 *>(tex + id * 8);
            const float4 a = __ldg(t), c = __ldg(t + 1);
            float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
            if (act != READ_TEXACT_NONE) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = tex_act(v[i], act);
            }
            if (LAYOUT == READ_FEAT_NCHW_F32) {
                float *o = static_cast<float *>(out) + b * 8 * hw + q;
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i * hw] = v[i];
            } else if (LAYOUT == READ_FEAT_NHWC_F32) {
                float4 *o = reinterpret_cast<float4 *>(static_cast<float *>(out) + p * 8);
                o[0] = make_float4(v[0], v[1], v[2], v[3]);
                o[1] = make_float4(v[4], v[5], v[6], v[7]);
            } else {
                __nv_bfloat162 r[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) r[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                *reinterpret_cast<uint4 *>(static_cast<__nv_bfloat16 *>(out) + p * 8) = *reinterpret_cast<uint4 *>(r);
            }
        } else {
            for (int c = 0; c < D; ++c) {
                const float v = tex_act(__ldg(tex + id * D + c), act);
                if (LAYOUT == READ_FEAT_NCHW_F32) static_cast<float *>(out)[(b * D + c) * hw + q] = v;
                else if (LAYOUT == READ_FEAT_NHWC_F32) static_cast<float *>(out)[p * D + c] = v;
                else static_cast<__nv_bfloat16 *>(out)[p * D + c] = __float2bfloat16_rn(v);
            }
        }
    }
}

// grad_tex[id,:] += grad_out[b,:,q].  Empty pixels (id 0) are pre-reduced per block in shared memory:
// in a sparse view millions of pixels would otherwise serialise on point 0's 8 addresses.
__global__ void gather_backward_kernel(const float *__restrict__ go, const float *__restrict__ ids, int B, int D, int h,
                                       int w, long long N, float *__restrict__ gt)
{
    extern __shared__ float zero_acc[];   // [D]
    for (int c = threadIdx.x; c < D; c += blockDim.x) zero_acc[c] = 0.f;
    __syncthreads();
    const long long hw = (long long)h * w;
    const long long total = (long long)B * hw;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        long long id = (long long)ids[p];
        if (id < 0) id = 0;
        if (id >= N) id = N - 1;
        const long long b = p / hw, q = p - b * hw;
        const float *g = go + b * D * hw + q;
        if (id == 0) {
            for (int c = 0; c < D; ++c) atomicAdd(&zero_acc[c], g[c * hw]);
        } else {
            for (int c = 0; c < D; ++c) atomicAdd(gt + id * D + c, g[c * hw]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x)
        if (zero_acc[c] != 0.f) atomicAdd(gt + c, zero_acc[c]);
}

static unsigned grid_for(long long total, int threads = 256)
{
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

template <int SRC>
static int launch_gather(const float *tex, int D, long long N, const void *src, int B, int h, int w, int layout,
                         int act, void *out, cudaStream_t st)
{
    const long long total = (long long)B * h * w;
    if (total == 0) return READ_OK;
    const unsigned g = grid_for(total);
    switch (layout) {
    case READ_FEAT_NCHW_F32:
        gather_kernel<SRC, READ_FEAT_NCHW_F32><<<g, 256, 0, st>>>(tex, D, N, src, B, h, w, act, out);
        break;
    case READ_FEAT_NHWC_F32:
        gather_kernel<SRC, READ_FEAT_NHWC_F32><<<g, 256, 0, st>>>(tex, D, N, src, B, h, w, act, out);
        break;
    case READ_FEAT_NHWC_BF16:
        gather_kernel<SRC, READ_FEAT_NHWC_BF16><<<g, 256, 0, st>>>(tex, D, N, src, B, h, w, act, out);
        break;
    default:
        set_error("gather: unknown layout %d", layout);
        return READ_ERR_INVALID;
    }
    RB_LAUNCH_CHECK();
    return READ_OK;
}

}  // namespace rb

using namespace rb;

extern "C" {

int read_texture_to_point_major(const float *tex_cn, int D, int64_t N, float *tex_nd, void *stream)
{
    RB_CHECK_ARG(tex_cn && tex_nd && D >= 1 && N >= 1, "texture transpose: bad arguments");
    RB_CHECK_ARG(D != 8 || (reinterpret_cast<uintptr_t>(tex_nd) & 15) == 0, "texture transpose: output must be 16B aligned");
    tex_to_point_major_kernel<<<grid_for(N), 256, 0, (cudaStream_t)stream>>>(tex_cn, D, N, tex_nd);
    RB_LAUNCH_CHECK();
    return READ_OK;
}

int read_texture_to_channel_major(const float *tex_nd, int D, int64_t N, float *tex_cn, void *stream)
{
    RB_CHECK_ARG(tex_cn && tex_nd && D >= 1 && N >= 1, "texture transpose: bad arguments");
    tex_to_channel_major_kernel<<<grid_for(N), 256, 0, (cudaStream_t)stream>>>(tex_nd, D, N, tex_cn);
    RB_LAUNCH_CHECK();
    return READ_OK;
}

static int check_gather(const float *tex, int D, int64_t N, const void *src, int B, int h, int w, const void *out)
{
    RB_CHECK_ARG(tex && src && out, "gather: null pointer");
    RB_CHECK_ARG(D >= 1 && N >= 1 && B >= 0 && h >= 0 && w >= 0, "gather: bad shape");
    RB_CHECK_ARG(D != 8 || (reinterpret_cast<uintptr_t>(tex) & 15) == 0, "gather: descriptors must be 16B aligned");
    RB_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "gather: output must be 16B aligned");
    return READ_OK;
}

int read_gather_from_index(const float *tex_nd, int D, int64_t N, const float *ids, int B, int h, int w, int layout,
                           int activation, void *out, void *stream)
{
    int rc = check_gather(tex_nd, D, N, ids, B, h, w, out);
    if (rc) return rc;
    return launch_gather<0>(tex_nd, D, N, ids, B, h, w, layout, activation, out, (cudaStream_t)stream);
}

int read_gather_from_zbuf(const float *tex_nd, int D, int64_t N, const uint64_t *zbuf_level, int B, int h, int w,
                          int layout, int activation, void *out, void *stream)
{
    int rc = check_gather(tex_nd, D, N, zbuf_level, B, h, w, out);
    if (rc) return rc;
    return launch_gather<1>(tex_nd, D, N, zbuf_level, B, h, w, layout, activation, out, (cudaStream_t)stream);
}

int read_gather_backward(const float *grad_out, const float *ids, int B, int D, int h, int w, int64_t N,
                         float *grad_tex_nd, void *stream)
{
    RB_CHECK_ARG(grad_out && ids && grad_tex_nd, "gather backward: null pointer");
    RB_CHECK_ARG(D >= 1 && D <= 1024 && N >= 1 && B >= 0 && h >= 0 && w >= 0, "gather backward: bad shape");
    const long long total = (long long)B * h * w;
    if (total == 0) return READ_OK;
    gather_backward_kernel<<<grid_for(total), 256, D * sizeof(float), (cudaStream_t)stream>>>(grad_out, ids, B, D, h, w,
                                                                                               N, grad_tex_nd);
    RB_LAUNCH_CHECK();
    return READ_OK;
}

}  // extern "C"
