// gemm_sm100.cuh -- host-side plan object of the tcgen05 GEMM / implicit-GEMM conv kernel (K7 + K4).
#pragma once
#include "common.cuh"

// Mirrors lb_gemm_desc of include/lb200.h (kept in sync by hand; plain C layout).
struct GemmDesc {
    const void* a0; int64_t a0_ld; int32_t a0_c;
    const void* a1; int64_t a1_ld; int32_t a1_c;
    int32_t B, H, W;
    int32_t taps;
    const void* w; int64_t w_ld;
    int32_t N;
    const void* bias;
    const void* bias2; int64_t bias2_ld;
    const void* res; int64_t res_ld;
    void* out; int64_t out_ld;
    int32_t mode;   // low byte: 0 linear epilogue, 1 GEGLU (N accumulators -> N/2 outputs); | LB_GEMM_STATIC_W | LB_GEMM_RELU
    // LayerNorm folded into this GEMM (see include/lb200.h)
    const void* ln_stats; int32_t ln_parts;
    const void* ln_csum; const void* ln_bias; float ln_eps;
    void* stats_out; int32_t stats_parts;
};

constexpr int kGemmMaxSegs = 12;

struct alignas(64) GemmParams {
    CUtensorMap tmA[2];
    CUtensorMap tmB;
    CUtensorMap tmBh;      // half-height box (BN/2 rows) for the 2-CTA multicast variant
    int num_segs;
    int seg_map[kGemmMaxSegs], seg_dy[kGemmMaxSegs], seg_dx[kGemmMaxSegs], seg_kb[kGemmMaxSegs];
    int total_kb;
    int W, H, B;
    int tw, th, tb;
    int tiles_x, tiles_y, tiles_m, tiles_n;
    int N;
    int mode;
    int stages;    // depth of the smem ring for this launch
    int static_w;  // weights may be fetched before griddepcontrol.wait (LB_GEMM_STATIC_W)
    __half* out; long long ldo;
    const __half* bias;
    const __half* bias2; long long bias2_ld;
    const __half* res; long long ldr;
    int* err_flag;
    int debug;     // profiling knobs (LB_GEMM_DEBUG): 1 = skip TMA issue, 2 = skip MMA issue
    int relu;      // mode 0: out = max(out, 0) (LB_GEMM_RELU)
    // LayerNorm fold: A holds the UN-normalised rows x; out = rstd*(acc - mu*csum[n]) + lnb[n] with (mu, rstd) from the
    // per-row partial sums the producing GEMM wrote (ln_stats[row][ln_parts] = (sum, sum of squares))
    const float2* ln_stats; int ln_parts; float ln_inv_k, ln_eps;
    const float* ln_csum; const float* ln_bias;
    float2* stats_out;     // [M][2*tiles_n]: (sum, sum of squares) of this launch's fp16 outputs per row and column part
};

struct GemmPlan {
    GemmParams p;
    int bn;        // N tile: 64 / 128 / 160 / 256
    int grid;
    int smem_bytes;
    int cluster;   // 1, or 2 = CTA pairs along M sharing the weight tile via TMA multicast
    int coresident;  // 1 = shallow-ring / single-accumulator variant, two CTAs per SM (single-wave problems)
};

int gemm_plan_build(lb_ctx* ctx, const GemmDesc& d, GemmPlan* plan);
int gemm_plan_launch(const GemmPlan& plan, cudaStream_t st);
int* lb_err_flag(lb_ctx* ctx);
