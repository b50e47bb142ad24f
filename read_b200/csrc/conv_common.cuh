// Shared pieces of the gated-conv kernels.
#pragma once
#include "common.cuh"

namespace rb {

// BasicConv tail (READ/models/unet.py:44-51): BN( A(f) * sigmoid(m) ), eval-mode BN folded to scale/shift.
// f, m already include their biases.
__device__ __forceinline__ float gated_epilogue(float f, float m, int elu, float scale, float shift)
{
    const float a = (elu && f <= 0.f) ? expm1f(f) : f;   // nn.ELU(alpha=1)
    const float s = 1.f / (1.f + expf(-m));              // nn.Sigmoid
    return fmaf(a * s, scale, shift);
}

// Fast variant for the bf16 tensor-core path: 1 MUFU for ELU (ex2), 1 MUFU for the gate (tanh).
__device__ __forceinline__ float gated_epilogue_fast(float f, float m, int elu, float scale, float shift)
{
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(f * 1.4426950408889634f));
    const float a = (elu && f <= 0.f) ? (e - 1.f) : f;
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * m));
    const float s = fmaf(0.5f, th, 0.5f);                // sigmoid(m) = 0.5*tanh(m/2)+0.5
    return fmaf(a * s, scale, shift);
}

// Same, activation chosen at compile time (the no-activation layers never touch the ex2 pipe).
template <bool ELU>
__device__ __forceinline__ float gate_fast(float f, float m, float scale, float shift)
{
    float a = f;
    if (ELU) {
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(f * 1.4426950408889634f));
        a = f <= 0.f ? (e - 1.f) : f;
    }
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * m));
    return fmaf(a * fmaf(0.5f, th, 0.5f), scale, shift);
}

// Same tail with the constants pre-folded by the caller: p = {bias_f, bias_m / 2, bn_scale / 2, bn_shift}.
//   sigmoid(m + b_m) = (1 + tanh(m/2 + b_m/2)) / 2   =>   y = a * (1 + th) * (scale / 2) + shift
// 9 instructions per output with ELU (FADD, FFMA, FMUL, 2 MUFU, FSETP, FADD, 2 FFMA), 5 without.
template <bool ELU>
__device__ __forceinline__ float gate_folded(float f, float m, const float4 p)
{
    const float v = f + p.x;
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(fmaf(m, 0.5f, p.y)));
    float a = v;
    if (ELU) {
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 1.4426950408889634f));
        a = v <= 0.f ? (e - 1.f) : v;
    }
    return fmaf(fmaf(a, th, a), p.z, p.w);
}

// floor(x / d) without integer division (a runtime IDIV costs ~20 instructions and every warp of every role decodes
// every tile).  floor((x + 0.5) * (1/d)) is exact for the x < 2^22 we ever see: the fractional part of (x+0.5)/d is at
// least 0.5/d away from an integer, far more than the fp32 rounding error.
__device__ __forceinline__ int fdiv_small(int x, float inv_d) { return (int)(((float)x + 0.5f) * inv_d); }

int generic_npad(int Cout);
int generic_kpad(int K);
int launch_generic(const read_conv_desc &d, cudaStream_t st);

// tcgen05 path (conv_tc.cu)
struct TcPlan;
bool tc_supported(const read_conv_desc &d);
int tc_plan_create(const read_conv_desc &d, TcPlan **out);
int tc_plan_launch(const TcPlan *p, cudaStream_t st, int max_ctas = 0);     // max_ctas > 0: persistent grid of at most that many CTAs
void tc_plan_destroy(TcPlan *p);
void tc_plan_set_reverse(TcPlan *p, int reverse);      // walk the tiles bottom-up (same result; L2 reuse between consecutive layers)


// tcgen05 CTA-pair path (conv_tc2.cu): cta_group::2 MMAs for the small-channel 3x3 layers
struct Tc2Plan;
bool tc2_supported(const read_conv_desc &d);
int tc2_plan_create(const read_conv_desc &d, Tc2Plan **out);
int tc2_plan_launch(const Tc2Plan *p, cudaStream_t st, int max_ctas = 0);
void tc2_plan_destroy(Tc2Plan *p);

// tcgen05 path with gathered A operand (conv_tc_gather.cu)
struct TcgPlan;
bool tcg_supported(const read_conv_desc &d);
int tcg_plan_create(const read_conv_desc &d, TcgPlan **out);
int tcg_plan_launch(const TcgPlan *p, cudaStream_t st, int max_ctas = 0);
void tcg_plan_destroy(TcgPlan *p);
int64_t tcg_weight_elems(int Cout, int Cin, int k);
int tcg_pack(const float *wf, const float *wm, int Cout, int Cin, int k, void *out, cudaStream_t st);

}  // namespace rb
