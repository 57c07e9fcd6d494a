// attn_sm100.cu -- K8: fused softmax(Q K^T / sqrt(d)) V for head dim 64 on tcgen05 (sm_100a).
//
// Replaces F.scaled_dot_product_attention under diffusers' AttnProcessor2_0 in
// every transformer block of the SDXL UNet (call site
// latentblending/diffusers_holder.py:336-344): self-attention (S = 4096 / 1024,
// Q,K,V slices of one fused-QKV activation) and cross-attention to the 77 text
// tokens (K,V slices of a [B,77,2C] projection).  fp16 in, fp32 softmax, fp16 out.
//
// One CTA = 256 query rows (two 128-row halves) of one (batch, head).  Per 128-key tile and half:
//   S = Q K^T   tcgen05.mma M128 N128 K64, Q/K K-major 128B-swizzled TMA tiles, S in TMEM
//   softmax     TWO threads per query row, each owning 64 of the tile's 128 key columns (16 softmax warps: two
//               groups of 8, one per half, ping-pong against the MMA issuer; the S -> softmax -> P V -> S chain of a
//               half is serial, so halving the softmax latency shortens every step): pass 1 row max (partners
//               exchange their partial maxima through smem + a 64-thread named barrier), pass 2
//               p = 2^(s*scale - m) -> fp16 P tile in smem (UMMA K-major layout).  The 64 S values stay in
//               registers between the passes: TMEM reads run at 64 B/clk/SM, so reading the fp32 S tile twice
//               (4096 clk per 256x128 tile) was 4x the tile's MMA time -- the limiter of the earlier versions
//   O += P V    tcgen05.mma M128 N64 K128 ACCUMULATING IN TMEM, V consumed MN-major from its TMA tile
// The running maximum is lazy: O (in TMEM) and the row sum are rescaled only when a row's maximum grows by more
// than 2^8 (then p <= 256, harmless in fp16/fp32); after the first tiles that almost never happens, so the
// softmax threads neither hold O in registers nor touch it per tile.  S(j+1) is issued behind P V(j) on the
// in-order tensor pipe, so observing S(j) complete implies P V(j-1) has completed: the rare rescale needs no
// extra barrier.  O is normalised by the row sum once, at the end.
// Bound: tensor pipe / MUFU.EX2 (16/clk/SM): at head dim 64 one exp feeds only 256 tensor FLOPs, so MUFU alone caps
// the tensor pipe at 50 %; every 4th pair of exps runs as a packed-fp32 polynomial on the FMA pipe instead.
// Algorithmic FLOPs = 4*Sq*Skv*64 per head.
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

using namespace sm100;

struct alignas(64) AttnParams {
    CUtensorMap tmQ, tmK, tmV;     // 3-D maps (columns, rows, batch), box (64, 128, 1), 128B swizzle
    int Sq, Skv, heads, B;
    int q_col0, k_col0, v_col0;    // column of head 0 inside each buffer
    __half* out; long long ldo;    // [B*Sq, heads*64]
    float scale_log2;              // softmax scale * log2(e)
    int* err_flag;
};

namespace {

constexpr int kD = 64, kBQ = 256, kBKV = 128;   // one CTA: 256 query rows (two 128-row halves), 128-key tiles
constexpr int kTileBytes = 128 * 64 * 2;        // 16 KiB: one Q half / K / V tile
constexpr int kPBytes = 128 * 128 * 2;          // 32 KiB: one P tile
constexpr int kKVStages = 2;
constexpr int kAttnSmem = (2 + 2 * kKVStages) * kTileBytes + 2 * kPBytes + 1024 + 256 + 4096 /*row exchange*/;
constexpr int kAttnThreads = 576;               // warp0 TMEM alloc + TMA, warp1 MMA, warps 2-9 / 10-17 softmax (<= 112 regs)
constexpr uint32_t kTmemCols = 512;             // S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384)

// Pipeline (per CTA, one KV tile j = one "step"):
//   tensor pipe:  S0[j] S1[j] | PV0[j] S0[j+1] | PV1[j] S1[j+1] | ...
//   softmax WG0:  -------- softmax(S0[j]) ------ | softmax(S0[j+1]) ...
//   softmax WG1:       -------- softmax(S1[j]) ------ | ...
// Each half's softmax overlaps the other half's MMAs; K/V tiles are loaded once per 256 query rows.
// kPolyMask: which of the 4 exp pairs of every 8-column group take the packed-fp32 polynomial 2^x on the FMA pipe instead
// of MUFU.EX2 (bit i = pair i).  0b1000 = every 4th pair (round 1); 0b1010 = every 2nd pair.
template <int kPolyMask>
__global__ void __launch_bounds__(kAttnThreads, 1) attn_tc_kernel(const __grid_constant__ AttnParams p) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t* sQ = smem;                                   // 2 halves
    uint8_t* sK = sQ + 2 * kTileBytes;                    // kKVStages
    uint8_t* sV = sK + kKVStages * kTileBytes;            // kKVStages
    uint8_t* sP = sV + kKVStages * kTileBytes;            // 2 halves
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;     // [2]
    uint64_t* k_empty = bars + 3;    // [2]
    uint64_t* v_full = bars + 5;     // [2]
    uint64_t* v_empty = bars + 7;    // [2]
    uint64_t* s_full = bars + 9;     // [2] per half
    uint64_t* p_ready = bars + 11;   // [2]
    uint64_t* o_final = bars + 13;   // [2] last P V of each half has completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
    float* xchg = reinterpret_cast<float*>(bars + 32);     // [parity 2][half 2][colpart 2][128 rows]

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kBQ, head = blockIdx.y, b = blockIdx.z;
    const int nkv = (p.Skv + kBKV - 1) / kBKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmQ);
        tma_prefetch_desc(&p.tmK);
        tma_prefetch_desc(&p.tmV);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_ready[i], 8);
            mbar_init(&o_final[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // ---------------- TMA producer (warp-uniform loop, one elected lane issues) ----------------
        if (elect_one()) {
            mbar_expect_tx(q_full, 2 * kTileBytes);
            tma_load_3d(sQ, &p.tmQ, q_full, p.q_col0 + head * kD, q0, b);
            tma_load_3d(sQ + kTileBytes, &p.tmQ, q_full, p.q_col0 + head * kD, q0 + 128, b);
        }
        __syncwarp();
        for (int j = 0; j < nkv; ++j) {
            const int st = j & 1;
            const uint32_t ph = (j >> 1) & 1;
            mbar_wait(&k_empty[st], ph ^ 1, p.err_flag, 11);
            if (elect_one()) {
                mbar_expect_tx(&k_full[st], kTileBytes);
                tma_load_3d(sK + st * kTileBytes, &p.tmK, &k_full[st], p.k_col0 + head * kD, j * kBKV, b);
            }
            __syncwarp();
            mbar_wait(&v_empty[st], ph ^ 1, p.err_flag, 12);
            if (elect_one()) {
                mbar_expect_tx(&v_full[st], kTileBytes);
                tma_load_3d(sV + st * kTileBytes, &p.tmV, &v_full[st], p.v_col0 + head * kD, j * kBKV, b);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (warp-uniform loop, one elected lane issues) ----------------
        constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc_f16(128, 64, 0, 1);   // B (= V) is MN-major
        const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP);
        auto issue_qk = [&](int half, int st, uint64_t* also_commit) {
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < kD / 16; ++k)
                    umma_f16(tmem_base + half * 128, make_smem_desc_sw128(q_addr + half * kTileBytes + k * 32, 16, 1024),
                             make_smem_desc_sw128(k_addr + st * kTileBytes + k * 32, 16, 1024), idesc_qk, k != 0);
                umma_commit(&s_full[half]);
                if (also_commit) umma_commit(also_commit);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int half, int st, bool first, bool last, uint64_t* also_commit) {
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < kBKV / 16; ++k)
                    umma_f16(tmem_base + 256 + half * 64,
                             make_smem_desc_sw128(p_addr + half * kPBytes + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                             make_smem_desc_sw128(v_addr + st * kTileBytes + k * 2048, 1024, 1024), idesc_pv,
                             (first && k == 0) ? 0u : 1u);      // O accumulates in TMEM across KV tiles
                if (last) umma_commit(&o_final[half]);
                if (also_commit) umma_commit(also_commit);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0, p.err_flag, 13);
        mbar_wait(&k_full[0], 0, p.err_flag, 14);
        tc_fence_after();
        issue_qk(0, 0, nullptr);
        issue_qk(1, 0, &k_empty[0]);
        for (int j = 0; j < nkv; ++j) {
            const int st = j & 1, stn = (j + 1) & 1;
            const uint32_t ph = j & 1, kvph = (j >> 1) & 1, kvphn = ((j + 1) >> 1) & 1;
            const bool more = j + 1 < nkv;
            mbar_wait(&v_full[st], kvph, p.err_flag, 16);
            mbar_wait(&p_ready[0], ph, p.err_flag, 15);
            tc_fence_after();
            issue_pv(0, st, j == 0, !more, nullptr);
            if (more) {
                mbar_wait(&k_full[stn], kvphn, p.err_flag, 14);
                tc_fence_after();
                issue_qk(0, stn, nullptr);
            }
            mbar_wait(&p_ready[1], ph, p.err_flag, 15);
            tc_fence_after();
            issue_pv(1, st, j == 0, !more, &v_empty[st]);
            if (more) issue_qk(1, stn, &k_empty[stn]);
        }
    } else if (warp >= 2) {
        // ---------------- softmax / output: two threads per query row, two independent halves ----------------
        // (any 4 consecutive warps cover the four TMEM lane quadrants: quadrant = warp % 4)
        const int half = (warp - 2) >> 3;
        const int colpart = ((warp - 2) >> 2) & 1;     // this thread's 64 key columns: [64*colpart, 64*colpart + 64)
        const int quad = warp & 3;                      // TMEM lane quadrant of this warp
        const int r = quad * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const uint32_t tmem_S = tmem_base + half * 128 + colpart * 64 + lane_off;
        const uint32_t tmem_O = tmem_base + 256 + half * 64 + colpart * 32 + lane_off;   // its 32 of the 64 O columns
        uint8_t* prow = sP + half * kPBytes + colpart * 16384 + r * 128;
        const uint32_t pair_bar = 1 + half * 4 + quad;                      // named barrier of the two partner warps
        float* x_mine = xchg + (half * 2 + colpart) * 128 + r;
        const float* x_other = xchg + (half * 2 + (colpart ^ 1)) * 128 + r;
        const int col0 = colpart * 64;
        float m_run = -INFINITY, l_run = 0.f;     // m_run in the scaled (log2) domain; l_run: this thread's columns only
        const float sc = p.scale_log2;
        int kv_valid = kBKV;
        // row max of this thread's 64 columns (TMEM loads double-buffered)
        auto max_pass = [&]() -> float {
            float mx = -INFINITY;
            uint32_t va[16], vb[16];
            tmem_ld_32x32b_x16(tmem_S, va);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t (&cur)[16] = (c & 1) ? vb : va;
                uint32_t (&nxt)[16] = (c & 1) ? va : vb;
                if (c < 3) tmem_ld_32x32b_x16(tmem_S + (c + 1) * 16, nxt);
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (kv_valid == kBKV || col0 + c * 16 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(cur[i]));
                if (c < 3) tmem_ld_wait();
            }
            return mx;
        };
        // p = 2^(s*sc - mref) -> fp16 -> smem (K-major, 128B swizzle: 16B chunk ^= row & 7), row sum and row max in
        // the SAME sweep over TMEM.  Packed fp32 FMAs; every 4th pair takes the polynomial 2^x so MUFU.EX2
        // (16/clk/SM) is not the only exp unit.
        auto exp_pass = [&](float mref, float& psum_out) -> float {
            float mx = -INFINITY;
            float2 psum2 = make_float2(0.f, 0.f);
            const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-mref, -mref);
            uint32_t va[16], vb[16];
            tmem_ld_32x32b_x16(tmem_S, va);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t (&cur)[16] = (c & 1) ? vb : va;
                uint32_t (&nxt)[16] = (c & 1) ? va : vb;
                if (c < 3) tmem_ld_32x32b_x16(tmem_S + (c + 1) * 16, nxt);
                if (kv_valid == kBKV) {
#pragma unroll
                    for (int i = 0; i < 16; i += 2) mx = fmax3(mx, __uint_as_float(cur[i]), __uint_as_float(cur[i + 1]));
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (col0 + c * 16 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(cur[i]));
                }
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint4 pk;
                    __half2* ph2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int col = col0 + c * 16 + g * 8 + 2 * i;
                        const float2 x2 = ffma2(make_float2(__uint_as_float(cur[g * 8 + 2 * i]),
                                                            __uint_as_float(cur[g * 8 + 2 * i + 1])), sc2, nm2);
                        float2 e2;
                        if ((kPolyMask >> i) & 1) {
                            e2 = ex2_poly2(x2);
                        } else {
                            e2.x = ex2_approx(x2.x);
                            e2.y = ex2_approx(x2.y);
                        }
                        if (kv_valid != kBKV) {
                            if (col >= kv_valid) e2.x = 0.f;
                            if (col + 1 >= kv_valid) e2.y = 0.f;
                        }
                        psum2 = fadd2(psum2, e2);
                        ph2[i] = __floats2half2_rn(e2.x, e2.y);
                    }
                    const int chunk = (c * 2 + g) ^ (r & 7);          // 16-byte chunk inside this row's 128 B
                    *reinterpret_cast<uint4*>(prow + chunk * 16) = pk;
                }
                if (c < 3) tmem_ld_wait();
            }
            psum_out = psum2.x + psum2.y;
            return mx;
        };
        for (int j = 0; j < nkv; ++j) {
            const uint32_t ph = j & 1;
            kv_valid = min(kBKV, p.Skv - j * kBKV);
            mbar_wait(&s_full[half], ph, p.err_flag, 17);
            tc_fence_after();
            // TMEM reads run at 64 B/clk/SM: reading the fp32 S tile twice (max sweep + exp sweep) costs 4096 clk per
            // 256x128 tile, 4x its MMAs.  So from the second tile on the exps are computed OPTIMISTICALLY against the
            // running reference maximum in the same sweep that finds the tile's maximum; only if a row outgrew the
            // reference by more than 2^8 (rare once the reference has settled) is the tile redone after rescaling.
            float psum = 0.f, mx;
            if (j == 0) mx = max_pass();
            else mx = exp_pass(m_run, psum);
            // partners exchange their partial maxima (double-buffered by step parity)
            x_mine[ph * 512] = mx;
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            mx = fmaxf(mx, x_other[ph * 512]);
            // (both partner warps see identical m_cand / m_run per lane, hence take the same decision)
            const float m_cand = mx * sc;
            if (__any_sync(0xffffffffu, m_cand - m_run > 8.0f)) {      // j == 0: m_run = -inf -> always
                const float m_new = fmaxf(m_run, m_cand);
                const float alpha = ex2_approx(m_run - m_new);          // first tile: 2^(-inf) = 0
                if (j > 0) {
                    // P V(j-1) has completed: S(j) was issued behind it on the in-order tensor pipe
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_O, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                    tmem_st_32x32b_x32(tmem_O, v);
                    tmem_st_wait();
                }
                l_run *= alpha;
                m_run = m_new;
                exp_pass(m_run, psum);                                  // redo this tile against the new reference
            }
            l_run += psum;
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[half]);
        }
        // ---- row sum: partners add their halves (fixed order), then output O / l for this thread's 32 O columns
        const int fp = (nkv & 1) * 512;           // the exchange buffer the last step did NOT use
        x_mine[fp] = l_run;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        const float l_tot = colpart ? (x_other[fp] + l_run) : (l_run + x_other[fp]);
        mbar_wait(&o_final[half], 0, p.err_flag, 19);
        tc_fence_after();
        const int q = q0 + half * 128 + r;
        const float inv = 1.0f / l_tot;
        __half* dst = p.out + ((long long)b * p.Sq + q) * p.ldo + head * kD + colpart * 32;
        {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_O, v);
            tmem_ld_wait();
            if (q < p.Sq) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 o;
                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        oh[i] = __floats2half2_rn(__uint_as_float(v[g * 8 + 2 * i]) * inv,
                                                  __uint_as_float(v[g * 8 + 2 * i + 1]) * inv);
                    *reinterpret_cast<uint4*>(dst + g * 8) = o;
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_rows_map(lb_ctx* ctx, CUtensorMap* m, const void* base, int64_t ld, int64_t cols, int rows, int B) {
    if (!ctx->tmap_encode) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        LB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres));
        LB_REQUIRE(f != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
        ctx->tmap_encode = f;
    }
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(ctx->tmap_encode);
    LB_REQUIRE(lb_aligned16(base) && ld % 8 == 0 && cols <= ld, "attention: operand alignment / stride");
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * rows};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(attention cols=%lld rows=%d B=%d) failed: %d",
               (long long)cols, rows, B, (int)r);
    return 0;
}

}  // namespace

int* lb_err_flag(lb_ctx* ctx);

struct AttnPlan {
    AttnParams p;
    dim3 grid;
};

int attn_plan_build(lb_ctx* ctx, const lb_attn_desc& d, AttnPlan* plan) {
    LB_REQUIRE(ctx && plan, "attention: null ctx/plan");
    LB_REQUIRE(d.q && d.k && d.v && d.out, "attention: null buffer");
    LB_REQUIRE(d.head_dim == 64, "attention: only head_dim 64 is implemented (got %d)", d.head_dim);
    LB_REQUIRE(d.B >= 1 && d.heads >= 1 && d.Sq >= 1 && d.Skv >= 1, "attention: bad sizes");
    LB_REQUIRE(d.out_ld % 8 == 0 && lb_aligned16(d.out), "attention: out alignment");
    AttnParams& p = plan->p;
    memset(&p, 0, sizeof(p));
    const int width = d.heads * 64;
    if (int e = encode_rows_map(ctx, &p.tmQ, d.q, d.q_ld, d.q_col0 + width, d.Sq, d.B)) return e;
    if (int e = encode_rows_map(ctx, &p.tmK, d.k, d.k_ld, d.k_col0 + width, d.Skv, d.B)) return e;
    if (int e = encode_rows_map(ctx, &p.tmV, d.v, d.v_ld, d.v_col0 + width, d.Skv, d.B)) return e;
    p.Sq = d.Sq; p.Skv = d.Skv; p.heads = d.heads; p.B = d.B;
    p.q_col0 = d.q_col0; p.k_col0 = d.k_col0; p.v_col0 = d.v_col0;
    p.out = static_cast<__half*>(d.out);
    p.ldo = d.out_ld;
    p.scale_log2 = d.scale * 1.4426950408889634f;
    p.err_flag = lb_err_flag(ctx);
    plan->grid = dim3((unsigned)lb_ceil_div(d.Sq, kBQ), (unsigned)d.heads, (unsigned)d.B);
    return 0;
}

template <int kPolyMask> static int attn_launch_variant(const AttnPlan& plan, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        LB_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<kPolyMask>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
        attr_set = true;
    }
    lb_launch_pdl(attn_tc_kernel<kPolyMask>, plan.grid, dim3(kAttnThreads), (size_t)kAttnSmem, st, plan.p);
    LB_LAUNCH_CHECK();
    return 0;
}

int attn_plan_launch(const AttnPlan& plan, cudaStream_t st) {
    static int poly = -1;
    if (poly < 0) poly = getenv("LB_ATTN_POLY") ? atoi(getenv("LB_ATTN_POLY")) : 8;
    switch (poly) {
        case 0: return attn_launch_variant<0>(plan, st);        // all exps on MUFU.EX2
        case 10: return attn_launch_variant<10>(plan, st);      // every 2nd pair on the FMA pipe
        case 14: return attn_launch_variant<14>(plan, st);      // three of four pairs on the FMA pipe
        default: return attn_launch_variant<8>(plan, st);       // every 4th pair (default)
    }
}

extern "C" int lb_attention(lb_ctx* ctx, const lb_attn_desc* desc, void* stream) {
    LB_REQUIRE(ctx && desc, "lb_attention: null argument");
    AttnPlan plan;
    if (int e = attn_plan_build(ctx, *desc, &plan)) return e;
    return attn_plan_launch(plan, lb_stream(stream));
}

// opaque handles for program.cu (AttnPlan holds CUtensorMaps and needs 64-byte alignment)
int attn_plan_build_opaque(lb_ctx* ctx, const lb_attn_desc& d, void** plan_out) {
    AttnPlan* plan = new AttnPlan();
    if (int e = attn_plan_build(ctx, d, plan)) {
        delete plan;
        return e;
    }
    *plan_out = plan;
    return 0;
}
int attn_plan_launch_opaque(void* plan, cudaStream_t st) { return attn_plan_launch(*static_cast<AttnPlan*>(plan), st); }
void attn_plan_free_opaque(void* plan) { delete static_cast<AttnPlan*>(plan); }
