// norm.cu -- K5 GroupNorm(+SiLU) and K6 LayerNorm on NHWC fp16 activations.
//
// Replace torch.nn.GroupNorm (+ F.silu) in the UNet resnets / transformer entry /
// conv_norm_out and torch.nn.LayerNorm in the 70 transformer blocks (call site
// latentblending/diffusers_holder.py:336-344).  HBM/L2-bound: GroupNorm reads x
// twice (stats pass, apply pass -- the second read is an L2 hit for every SDXL
// activation) and writes once; LayerNorm is single-read (row kept in registers).
// Deterministic: fixed-order reductions, no atomics.  fp32 statistics, fp64 final
// combine; output rounded to fp16 after the affine and again after SiLU, like the
// reference's two separate torch ops.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxC = 2560;

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

// ---- GroupNorm pass 1: per (batch, row-chunk) partial sums per group ------------------
__global__ void __launch_bounds__(kThreads)
gn_partial_kernel(const __half* __restrict__ x, long long ld, int C, int HW, int groups, int rows_per_chunk,
                  float2* __restrict__ partial /*[B][chunks][groups]*/) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
    const int row0 = chunk * rows_per_chunk;
    const int row1 = min(HW, row0 + rows_per_chunk);
    const int pairs = C >> 1;
    __shared__ float2 s_part[kMaxC / 2];
    const __half* base = x + ((long long)b * HW) * ld;
    for (int c2 = threadIdx.x; c2 < pairs; c2 += kThreads) {
        float s = 0.f, q = 0.f;
        for (int r = row0; r < row1; ++r) {
            const float2 v = __half22float2(*reinterpret_cast<const __half2*>(base + (long long)r * ld + 2 * c2));
            s += v.x + v.y;
            q = fmaf(v.x, v.x, q);
            q = fmaf(v.y, v.y, q);
        }
        s_part[c2] = make_float2(s, q);
    }
    __syncthreads();
    if (threadIdx.x < groups) {
        const int g = threadIdx.x;
        const int ppg = pairs / groups;       // channel pairs per group (channels/group is even)
        float s = 0.f, q = 0.f;
        for (int i = 0; i < ppg; ++i) {
            const float2 v = s_part[g * ppg + i];
            s += v.x;
            q += v.y;
        }
        partial[((long long)b * chunks + chunk) * groups + g] = make_float2(s, q);
    }
}

// ---- GroupNorm pass 2: finish statistics, normalise, affine, optional SiLU ----------------
__global__ void __launch_bounds__(kThreads)
gn_apply_kernel(const __half* __restrict__ x, long long ld, int C, int HW, int groups, int chunks,
                const float2* __restrict__ partial, const __half* __restrict__ gamma,
                const __half* __restrict__ beta, float eps, int do_silu, __half* __restrict__ out, long long ldo,
                int rows_per_block) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y;
    __shared__ float s_mean[64], s_rstd[64];
    __shared__ float s_scale[kMaxC], s_shift[kMaxC];
    const int cpg = C / groups;
    if (threadIdx.x < groups) {
        double s = 0.0, q = 0.0;
        for (int c = 0; c < chunks; ++c) {
            const float2 v = partial[((long long)b * chunks + c) * groups + threadIdx.x];
            s += v.x;
            q += v.y;
        }
        const double n = (double)HW * cpg;
        const double mean = s / n;
        double var = q / n - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = (float)mean;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kThreads) {
        const int g = c / cpg;
        const float sc = s_rstd[g] * __half2float(gamma[c]);
        s_scale[c] = sc;
        s_shift[c] = __half2float(beta[c]) - s_mean[g] * sc;
    }
    __syncthreads();
    const int vecs = C >> 3;
    const int row0 = blockIdx.x * rows_per_block;
    const int row1 = min(HW, row0 + rows_per_block);
    const long long total = (long long)(row1 - row0) * vecs;
    for (long long i = threadIdx.x; i < total; i += kThreads) {
        const int r = row0 + (int)(i / vecs);
        const int c0 = (int)(i % vecs) * 8;
        const long long row = (long long)b * HW + r;
        const uint4 v = *reinterpret_cast<const uint4*>(x + row * ld + c0);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = __half22float2(h[j]);
            float y0 = lb_round_h(fmaf(f.x, s_scale[c0 + 2 * j], s_shift[c0 + 2 * j]));
            float y1 = lb_round_h(fmaf(f.y, s_scale[c0 + 2 * j + 1], s_shift[c0 + 2 * j + 1]));
            if (do_silu) {
                y0 = silu_f(y0);
                y1 = silu_f(y1);
            }
            oh[j] = __floats2half2_rn(y0, y1);
        }
        *reinterpret_cast<uint4*>(out + row * ldo + c0) = o;
    }
}

// ---- LayerNorm: one warp per row, row held in registers -------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(kThreads)
ln_kernel(const __half* __restrict__ x, long long ld, long long rows, int C, const __half* __restrict__ gamma,
          const __half* __restrict__ beta, float eps, __half* __restrict__ out, long long ldo) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (kThreads / 32) + warp;
    if (row >= rows) return;
    const int vecs = C >> 3;
    uint4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vi = lane + i * 32;
        if (vi < vecs) {
            v[i] = *reinterpret_cast<const uint4*>(x + row * ld + vi * 8);
            const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                s += f.x + f.y;
            }
        }
    }
    s = lb_warp_sum(s);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vi = lane + i * 32;
        if (vi < vecs) {
            const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                q = fmaf(f.x - mean, f.x - mean, q);
                q = fmaf(f.y - mean, f.y - mean, q);
            }
        }
    }
    q = lb_warp_sum(q);
    const float rstd = rsqrtf(q / (float)C + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int vi = lane + i * 32;
        if (vi < vecs) {
            const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
            const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gamma + vi * 8));
            const uint4 bv = __ldg(reinterpret_cast<const uint4*>(beta + vi * 8));
            const __half2* gh = reinterpret_cast<const __half2*>(&gv);
            const __half2* bh = reinterpret_cast<const __half2*>(&bv);
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                const float2 g = __half22float2(gh[j]);
                const float2 bb = __half22float2(bh[j]);
                oh[j] = __floats2half2_rn(fmaf((f.x - mean) * rstd, g.x, bb.x), fmaf((f.y - mean) * rstd, g.y, bb.y));
            }
            *reinterpret_cast<uint4*>(out + row * ldo + vi * 8) = o;
        }
    }
}

}  // namespace

static int gn_chunks(int B, int HW, int sm_count) {
    int chunks = (2 * sm_count + B - 1) / B;
    if (chunks > HW) chunks = HW;
    if (chunks > 256) chunks = 256;
    if (chunks < 1) chunks = 1;
    return chunks;
}

extern "C" size_t lb_groupnorm_workspace_bytes(lb_ctx* ctx, int B, int HW, int groups) {
    if (!ctx) return 0;
    return (size_t)B * gn_chunks(B, HW, ctx->sm_count) * groups * sizeof(float2);
}

extern "C" int lb_groupnorm(lb_ctx* ctx, const void* x, int64_t ld, int B, int HW, int C, int groups,
                            const void* gamma, const void* beta, float eps, int silu, void* out, int64_t ldo,
                            void* workspace, void* stream) {
    LB_REQUIRE(ctx && x && gamma && beta && out && workspace, "lb_groupnorm: null argument");
    LB_REQUIRE(groups >= 1 && groups <= 64 && C % groups == 0 && (C / groups) % 2 == 0,
               "lb_groupnorm: channels per group must be even (C=%d groups=%d)", C, groups);
    LB_REQUIRE(C % 8 == 0 && C <= kMaxC, "lb_groupnorm: C must be a multiple of 8 and <= %d (got %d)", kMaxC, C);
    LB_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && lb_aligned16(x) && lb_aligned16(out), "lb_groupnorm: alignment");
    const int chunks = gn_chunks(B, HW, ctx->sm_count);
    const int rpc = (int)lb_ceil_div(HW, chunks);
    const int used_chunks = (int)lb_ceil_div(HW, rpc);
    cudaStream_t st = lb_stream(stream);
    lb_launch_pdl(gn_partial_kernel, dim3(used_chunks, B), kThreads, 0, st, (const __half*)x, ld, C, HW, groups, rpc,
                                                                (float2*)workspace);
    LB_LAUNCH_CHECK();
    // apply: ~4 blocks per SM
    int blocks = (4 * ctx->sm_count + B - 1) / B;
    if (blocks > HW) blocks = HW;
    const int rpb = (int)lb_ceil_div(HW, blocks);
    blocks = (int)lb_ceil_div(HW, rpb);
    lb_launch_pdl(gn_apply_kernel, dim3(blocks, B), kThreads, 0, st, (const __half*)x, ld, C, HW, groups, used_chunks,
                                                          (const float2*)workspace, (const __half*)gamma,
                                                          (const __half*)beta, eps, silu, (__half*)out, ldo, rpb);
    LB_LAUNCH_CHECK();
    return 0;
}

extern "C" int lb_layernorm(lb_ctx* ctx, const void* x, int64_t ld, int64_t rows, int C, const void* gamma,
                            const void* beta, float eps, void* out, int64_t ldo, void* stream) {
    LB_REQUIRE(ctx && x && gamma && beta && out, "lb_layernorm: null argument");
    LB_REQUIRE(C % 8 == 0 && C <= 2048, "lb_layernorm: C must be a multiple of 8 and <= 2048 (got %d)", C);
    LB_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && lb_aligned16(x) && lb_aligned16(out) && lb_aligned16(gamma) &&
                   lb_aligned16(beta), "lb_layernorm: alignment");
    if (rows == 0) return 0;
    const unsigned grid = (unsigned)lb_ceil_div(rows, kThreads / 32);
    cudaStream_t st = lb_stream(stream);
    const int vecs = C / 8;
    if (vecs <= 64)
        lb_launch_pdl(ln_kernel<2>, grid, kThreads, 0, st, (const __half*)x, ld, rows, C, (const __half*)gamma,
                                                (const __half*)beta, eps, (__half*)out, ldo);
    else if (vecs <= 160)
        lb_launch_pdl(ln_kernel<5>, grid, kThreads, 0, st, (const __half*)x, ld, rows, C, (const __half*)gamma,
                                                (const __half*)beta, eps, (__half*)out, ldo);
    else
        lb_launch_pdl(ln_kernel<8>, grid, kThreads, 0, st, (const __half*)x, ld, rows, C, (const __half*)gamma,
                                                (const __half*)beta, eps, (__half*)out, ldo);
    LB_LAUNCH_CHECK();
    return 0;
}
