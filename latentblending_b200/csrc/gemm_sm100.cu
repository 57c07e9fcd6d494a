// gemm_sm100.cu -- K7/K4: tcgen05 GEMM and implicit-GEMM convolution for sm_100a.
//
//   out[M, N] = epilogue( sum_seg A_seg[M, K_seg] * W[N, K]^T )
//
// Replaces every dense contraction of the SDXL UNet the reference runs through
// cuBLAS / cuDNN (call site latentblending/diffusers_holder.py:336-344): the
// Linear layers (to_q/k/v, to_out, proj_in/out, GEGLU FF), the 3x3 / 1x1
// convolutions of the resnets and samplers (implicit GEMM: one K-segment per
// filter tap, the A tile of a tap is a TMA box of the NHWC activation shifted
// by (dy,dx) with hardware zero fill at the borders -- no im2col buffer), and
// the resnet shortcut folded in as an extra K-segment from a second tensor.
//
// Structure (one CTA per SM, persistent over output tiles, 576 threads):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor 4D (A) / 2D (W) into a
//              STAGES-deep 128B-swizzled smem ring, mbarrier full/empty pairs
//   warp 1   : MMA issuer    -- one thread issues tcgen05.mma (M=128, N=BN,
//              K=16, fp16 in / fp32 accumulate in TMEM), tcgen05.commit frees
//              smem slots and publishes the accumulator
//   warps 2-17: epilogue     -- tcgen05.ld TMEM->registers, + bias / per-batch
//              bias (time embedding) / residual, or GEGLU, fp16 store; FOUR warps
//              per TMEM lane quadrant split the tile's 16-column chunks (r02r/r02s: with
//              two warps per quadrant the epilogue was latency-bound at 0.39 IPC and the
//              exposed drain of the single-tile launches cost ~1.4 ms per UNet forward);
//              residual rows are prefetched while the accumulator is still being
//              produced; double-buffered accumulators overlap the epilogue with the
//              next tile's MMAs
// Weight (B operand) tiles of the first pipeline stages are requested BEFORE
// griddepcontrol.wait when the caller marks the weights static (LB_GEMM_STATIC_W):
// their HBM latency hides behind the tail of the previous kernel.
// Bound: tensor pipe; algorithmic FLOPs = 2*M*N*K.
#include "gemm_sm100.cuh"
#include <stdlib.h>

#include "sm100.cuh"

using namespace sm100;

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kEpiParts = 4;                 // epilogue warps per TMEM lane quadrant (they split the tile's columns)
constexpr int kEpiWarps = 4 * kEpiParts;     // 16
constexpr int kThreadsGemm = 64 + 32 * kEpiWarps;   // warp 0 TMA, warp 1 MMA, warps 2.. epilogue
constexpr int kABytes = kBM * kBK * 2;  // 16 KiB

template <int BN> struct Cfg {
    static constexpr int b_bytes = BN * kBK * 2;
    static constexpr int stage_bytes = kABytes + b_bytes;
    // ring depth is a per-launch parameter (GemmParams::stages): `deep` = as many stages as 227 KB of shared memory
    // allow (one k-block is 256-320 clk of MMA, a TMA round trip ~1900 clk; measured +4-11 % on the long-K
    // convolutions), `stages` = the shallower ring the short-K problems were tuned with
    static constexpr int stages = (BN <= 64) ? 8 : (BN <= 128) ? 6 : (BN <= 160) ? 5 : 4;
    static constexpr int deep = (BN <= 64) ? 8 : (BN <= 128) ? 7 : (BN <= 160) ? 6 : 4;
    static constexpr int tmem_cols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                   : (2 * BN <= 256) ? 256 : 512;
    static constexpr int tmem_cols_single = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : (BN <= 256) ? 256 : 512;
    // co-resident variant: ring depth such that two CTAs (+ 1 KB reserved each) fit in the SM's 228 KB
    static constexpr int cr_stages = (BN <= 64) ? 4 : 3;
    static constexpr int smem_bytes(int nst) { return nst * stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/; }
    static constexpr int pair_stage_bytes = kABytes + b_bytes / 2;   // bytes ONE CTA of a pair stages per k-block
};

// exact-erf GELU, 0.5 x (1 + erf(x / sqrt 2)), with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below
// the fp16 rounding the reference applies to gelu(gate)): ~14 FMA-pipe instructions + MUFU.RCP + MUFU.EX2 instead
// of erff's two-branch polynomial -- the GEGLU epilogue of the FF-in GEMMs was issue-bound next to a K = 1280 main loop.
// For z < 0, 1 + erf(z) = erfc(|z|) is formed directly (no cancellation in the negative tail).
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float erfc_abs = poly * t * exp2f(-1.4426950408889634f * z * z);     // erfc(|z|)
    const float one_plus_erf = x >= 0.f ? 2.0f - erfc_abs : erfc_abs;
    return 0.5f * x * one_plus_erf;
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __half22float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// (mu, rstd) of one row of the LayerNorm-folded A operand from the producer's per-row partial sums (fixed order).
// The partials of a row are contiguous (<= 64 x float2): they are fetched as float4 pairs, eight loads in flight at
// a time -- a scalar loop over them serialises ~16 L2 round trips (10k clk per tile: slower than the LayerNorm
// launches the fold removes; measured in r02a).
__device__ __forceinline__ void ln_row_stats(const GemmParams& p, long long row, bool ok, float& mu, float& rstd) {
    mu = 0.f;
    rstd = 1.f;
    if (!ok) return;
    float s = 0.f, q = 0.f;
    const float4* st = reinterpret_cast<const float4*>(p.ln_stats + row * p.ln_parts);   // ln_parts is even
    const int pairs = p.ln_parts >> 1;
    for (int base = 0; base < pairs; base += 8) {
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (base + i < pairs) ? __ldcg(st + base + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s += v[i].x;
            q += v[i].y;
            s += v[i].z;
            q += v[i].w;
        }
    }
    mu = s * p.ln_inv_k;
    double var = (double)q * (double)p.ln_inv_k - (double)mu * (double)mu;
    if (var < 0.0) var = 0.0;
    rstd = rsqrtf((float)var + p.ln_eps);
}

// CL = 2 (cta_group::2): a CTA PAIR owns a 256 x BN tile -- two vertically adjacent 128-row M tiles of one N tile.
// Each CTA stages its own A tile and HALF of the weight tile; the leader issues tcgen05.mma.cta_group::2 (M = 256),
// each SM's tensor core accumulates its 128 rows in its own TMEM and the weight halves are shared across the pair,
// so per-SM shared-memory traffic (the limiter of single-CTA M=128 MMAs on Blackwell) drops by BN/2 rows per k-block.
// CR = 1: the CO-RESIDENT variant for single-wave problems (every CTA owns one tile: the transformer's 2048-row
// linears).  Shallow smem ring (<= 110 KB) + ONE accumulator (<= 256 TMEM columns) + <= 96 registers, so two CTAs fit
// on an SM: under PDL the NEXT kernel's CTAs become resident while this kernel is still running, and their prologue
// (barrier init, TMEM allocation, descriptor prefetch, the first weight tiles) no longer waits for this kernel's CTAs
// to exit -- with one 200 KB CTA per SM that hand-over (~2 us of a ~10 us kernel) was fully exposed.
template <int BN, int CL, int CR>
__global__ void __launch_bounds__(kThreadsGemm, CR ? 2 : 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
    using C = Cfg<BN>;
    constexpr int kAcc = CR ? 1 : 2;            // accumulator buffers in TMEM
    constexpr int kTmemCols = CR ? C::tmem_cols_single : C::tmem_cols;
    pdl_launch_dependents();       // the next kernel may start its launch + prologue while this one runs
    const uint32_t crank = (CL == 2) ? cluster_ctarank() : 0u;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // SWIZZLE_128B needs 1024 B alignment
    uint8_t* smem_a = smem;
    const int nst = p.stages;      // ring depth of this launch (Cfg::stages or Cfg::deep)
    uint8_t* smem_b = smem + nst * kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nst * C::stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + nst;
    uint64_t* tmem_full = bars + 2 * nst;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA[0]);
        tma_prefetch_desc(&p.tmA[1]);
        tma_prefetch_desc(&p.tmB);
        for (int s = 0; s < nst; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], kEpiWarps * CL);   // pair: the leader waits for both CTAs' epilogue warps
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        if (CL == 2) tmem_alloc_2sm(tmem_slot, kTmemCols);
        else tmem_alloc(tmem_slot, kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();          // peer barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile index space: (m-group, n) with CL vertically adjacent M tiles per group
    const int m_groups = (p.tiles_m + CL - 1) / CL;
    const int num_tiles = m_groups * p.tiles_n;
    const int tile0 = (CL == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = (CL == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    // Static weights do not depend on the previous grid: request the B tiles of the first ring stages now, while
    // that grid is still finishing (the A tiles of the same stages follow after griddepcontrol.wait).
    int early_kb = 0;
    if (CL == 1 && p.static_w && !(p.debug & 1) && tile0 < num_tiles)
        early_kb = p.total_kb < nst ? p.total_kb : nst;
    if (warp == 0 && early_kb > 0) {
        if (elect_one()) {
            const int n_tile = tile0 / m_groups;
            for (int kb = 0; kb < early_kb; ++kb) {
                mbar_expect_tx(&full[kb], C::stage_bytes);
                tma_load_2d(smem_b + kb * C::b_bytes, &p.tmB, &full[kb], kb * kBK, n_tile * BN);
            }
        }
        __syncwarp();
    }
    pdl_wait();                               // everything above overlapped the previous kernel; its outputs are visible now

    if (warp == 0) {
        // ===================== TMA producer (whole warp runs the loop, one elected lane issues) =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            const int m_tile = (tile % m_groups) * CL + (int)crank, n_tile = tile / m_groups;
            const int x0 = (m_tile % p.tiles_x) * p.tw;
            const int y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.th;
            const int b0 = (m_tile / (p.tiles_x * p.tiles_y)) * p.tb;   // m_tile == tiles_m (odd tail): fully OOB -> zeros
            int kb_global = 0;
            for (int s = 0; s < p.num_segs; ++s) {
                const CUtensorMap* ma = &p.tmA[p.seg_map[s]];
                const int dy = p.seg_dy[s], dx = p.seg_dx[s];
                for (int kb = 0; kb < p.seg_kb[s]; ++kb, ++kb_global) {
                    mbar_wait(&empty[stage], phase ^ 1, p.err_flag, 1);
                    if (elect_one()) {
                    if (p.debug & 1) {                 // profiling: no data movement, just hand the slot over
                        if (CL == 1 || crank == 0) mbar_arrive(&full[stage]);
                    } else if (CL == 2) {
                        // both CTAs' bytes are counted on the LEADER's full barrier
                        const uint32_t lead_full = map_to_cta(&full[stage], 0);
                        if (crank == 0) mbar_expect_tx(&full[stage], 2 * C::pair_stage_bytes);
                        tma_load_4d_2sm(smem_a + stage * kABytes, ma, lead_full, kb * kBK, x0 + dx, y0 + dy, b0);
                        tma_load_2d_2sm(smem_b + stage * C::b_bytes, &p.tmBh, lead_full, kb_global * kBK,
                                        n_tile * BN + (int)crank * (BN / 2));
                    } else if (tile == tile0 && kb_global < early_kb) {
                        // expect_tx and the weight tile were issued before griddepcontrol.wait
                        tma_load_4d(smem_a + stage * kABytes, ma, &full[stage], kb * kBK, x0 + dx, y0 + dy, b0);
                    } else {
                        mbar_expect_tx(&full[stage], C::stage_bytes);
                        tma_load_4d(smem_a + stage * kABytes, ma, &full[stage], kb * kBK, x0 + dx, y0 + dy, b0);
                        tma_load_2d(smem_b + stage * C::b_bytes, &p.tmB, &full[stage], kb_global * kBK, n_tile * BN);
                    }
                    }
                    __syncwarp();
                    if (++stage == nst) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && crank == 0) {
        // ===================== MMA issuer (pair: leader CTA only) =====================
        constexpr uint32_t idesc = make_idesc_f16(kBM * CL, BN);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1, p.err_flag, 2);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < p.total_kb; ++kb) {
                mbar_wait(&full[stage], phase, p.err_flag, 3);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem_a + stage * kABytes);
                const uint32_t b_addr = smem_u32(smem_b + stage * C::b_bytes);
                if (elect_one()) {
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                    const uint64_t adesc = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                    if (p.debug & 2) continue;         // profiling: no tensor work
                    if (CL == 2) umma_f16_2sm(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
                    else umma_f16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
                }
                if (CL == 2) umma_commit_2sm(&empty[stage], (uint16_t)0x3);   // frees the slot in both CTAs
                else umma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == nst) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) {
                if (CL == 2) umma_commit_2sm(&tmem_full[acc], (uint16_t)0x3);
                else umma_commit(&tmem_full[acc]);
            }
            __syncwarp();
            if (++acc == kAcc) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 2) {
        // ===================== epilogue (16 warps: four per TMEM lane quadrant, splitting the columns) =====================
        // r02r (ncu): the GEGLU epilogue of the FF-in GEMM took ~9 900 clk per tile with 8 warps (~1 900 instructions per
        // thread at 0.39 IPC per scheduler: two-and-a-half resident warps do not hide the FMA / TMEM latencies) against
        // 5 120 clk of MMAs -- the tensor pipe idled half the time waiting for a free accumulator.  16 warps with 16-column
        // chunks halve the per-warp work of every epilogue (and the exposed drain of the single-tile launches).
        const int q = warp & 3;                 // TMEM lane quadrant this warp may access
        const int part = (warp - 2) >> 2;       // 0..3: which quarter of the tile's 16-column chunks
        const int r = q * 32 + lane;            // accumulator row inside the tile
        const int ww = r % p.tw, hh = (r / p.tw) % p.th, bb = r / (p.tw * p.th);
        const int mode = p.mode;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            const int m_tile = (tile % m_groups) * CL + (int)crank, n_tile = tile / m_groups;
            const int x = (m_tile % p.tiles_x) * p.tw + ww;
            const int y = ((m_tile / p.tiles_x) % p.tiles_y) * p.th + hh;
            const int b = (m_tile / (p.tiles_x * p.tiles_y)) * p.tb + bb;
            const bool row_ok = (x < p.W) && (y < p.H) && (b < p.B) && (m_tile < p.tiles_m);
            const long long row = ((long long)b * p.H + y) * p.W + x;
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            if (mode == 0) {
                constexpr int NCH = BN / 16;                      // 16-column chunks of the tile
                constexpr int NPER = (NCH + kEpiParts - 1) / kEpiParts;   // chunks per warp (the last part may get fewer)
                const int c_begin = part * NPER;
                const int c_end = (c_begin + NPER < NCH) ? c_begin + NPER : NCH;
                const int n_base = n_tile * BN;
                const __half* res_row = p.res ? p.res + row * p.ldr : nullptr;
                // residual of the first chunk: requested BEFORE the accumulator is complete (overlaps the MMAs)
                uint4 rnext[2];
                auto load_res = [&](int c, uint4 (&dst)[2]) {
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const int n = n_base + c * 16 + g * 8;
                        dst[g] = (res_row && row_ok && n < p.N) ? *reinterpret_cast<const uint4*>(res_row + n)
                                                               : make_uint4(0, 0, 0, 0);
                    }
                };
                if (c_begin < c_end) load_res(c_begin, rnext);
                float ln_mu = 0.f, ln_rstd = 1.f;
                if (p.ln_stats) ln_row_stats(p, row, row_ok, ln_mu, ln_rstd);   // overlaps the MMAs of this tile
                float st_sum = 0.f, st_sq = 0.f;
                mbar_wait(&tmem_full[acc], acc_phase, p.err_flag, 4);
                tc_fence_after();
#pragma unroll
                for (int i = 0; i < NPER; ++i) {
                    const int c = c_begin + i;
                    if (c < c_end) {                    // warp-uniform
                        uint32_t v[16];
                        tmem_ld_32x32b_x16(t_addr + c * 16, v);
                        uint4 rcur[2];
#pragma unroll
                        for (int g = 0; g < 2; ++g) rcur[g] = rnext[g];
                        if (c + 1 < c_end) load_res(c + 1, rnext);   // in flight while this chunk is processed
                        tmem_ld_wait();
                        const int n0 = n_base + c * 16;
                        if (row_ok && n0 < p.N) {
#pragma unroll
                            for (int g = 0; g < 2; ++g) {
                                const int n = n0 + g * 8;
                                if (n < p.N) {      // N is a multiple of 8 (checked on the host)
                                    float acc8[8];
#pragma unroll
                                    for (int j = 0; j < 8; ++j) acc8[j] = __uint_as_float(v[g * 8 + j]);
                                    float t8[8];
                                    if (p.ln_stats) {
                                        float c8[8];
                                        load8f(p.ln_csum + n, c8);
                                        load8f(p.ln_bias + n, t8);
#pragma unroll
                                        for (int j = 0; j < 8; ++j)
                                            acc8[j] = fmaf(ln_rstd, acc8[j] - ln_mu * c8[j], t8[j]);
                                    } else if (p.bias) {
                                        unpack8(__ldg(reinterpret_cast<const uint4*>(p.bias + n)), t8);
#pragma unroll
                                        for (int j = 0; j < 8; ++j) acc8[j] += t8[j];
                                    }
                                    if (p.bias2) {
                                        unpack8(__ldg(reinterpret_cast<const uint4*>(p.bias2 + (long long)b * p.bias2_ld + n)), t8);
#pragma unroll
                                        for (int j = 0; j < 8; ++j) acc8[j] += t8[j];
                                    }
                                    if (p.res) {
                                        unpack8(rcur[g], t8);
#pragma unroll
                                        for (int j = 0; j < 8; ++j) acc8[j] += t8[j];
                                    }
                                    if (p.relu) {
#pragma unroll
                                        for (int j = 0; j < 8; ++j) acc8[j] = fmaxf(acc8[j], 0.f);
                                    }
                                    uint4 o;
                                    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
                                    for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(acc8[2 * j], acc8[2 * j + 1]);
                                    *reinterpret_cast<uint4*>(p.out + row * p.ldo + n) = o;
                                    if (p.stats_out) {       // statistics of the STORED (fp16-rounded) values
#pragma unroll
                                        for (int j = 0; j < 4; ++j) {
                                            const float2 f = __half22float2(oh[j]);
                                            st_sum += f.x + f.y;
                                            st_sq = fmaf(f.x, f.x, fmaf(f.y, f.y, st_sq));
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
                if (p.stats_out && row_ok)
                    p.stats_out[row * (kEpiParts * p.tiles_n) + kEpiParts * n_tile + part] = make_float2(st_sum, st_sq);
            } else {
                // GEGLU: tile columns [0,BN/2) are "value", [BN/2,BN) the matching "gate" (weights are
                // row-interleaved per tile on the host); out = (v+bv) * gelu(g+bg), BN/2 outputs per tile.
                constexpr int HN = BN / 2;
                constexpr int NCH = HN / 16;
                constexpr int NPER = (NCH + kEpiParts - 1) / kEpiParts;
                const int c_begin = part * NPER;
                const int c_end = (c_begin + NPER < NCH) ? c_begin + NPER : NCH;
                const int o_base = n_tile * HN;        // output column base
                const int a_base = n_tile * BN;        // accumulator (bias) column base
                float ln_mu = 0.f, ln_rstd = 1.f;
                if (p.ln_stats) ln_row_stats(p, row, row_ok, ln_mu, ln_rstd);
                mbar_wait(&tmem_full[acc], acc_phase, p.err_flag, 4);
                tc_fence_after();
#pragma unroll 1
                for (int c = c_begin; c < c_end; ++c) {
                    uint32_t vv[16], vg[16];
                    tmem_ld_32x32b_x16(t_addr + c * 16, vv);
                    tmem_ld_32x32b_x16(t_addr + HN + c * 16, vg);
                    tmem_ld_wait();
                    if (row_ok) {
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            float bv[8], bg[8];
                            const int jn = c * 16 + g * 8;
                            if (p.ln_stats) {
                                float cv[8], cg[8];
                                load8f(p.ln_csum + a_base + jn, cv);
                                load8f(p.ln_csum + a_base + HN + jn, cg);
                                load8f(p.ln_bias + a_base + jn, bv);
                                load8f(p.ln_bias + a_base + HN + jn, bg);
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    vv[g * 8 + j] = __float_as_uint(ln_rstd * (__uint_as_float(vv[g * 8 + j]) - ln_mu * cv[j]));
                                    vg[g * 8 + j] = __float_as_uint(ln_rstd * (__uint_as_float(vg[g * 8 + j]) - ln_mu * cg[j]));
                                }
                            } else if (p.bias) {
                                unpack8(__ldg(reinterpret_cast<const uint4*>(p.bias + a_base + jn)), bv);
                                unpack8(__ldg(reinterpret_cast<const uint4*>(p.bias + a_base + HN + jn)), bg);
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) bv[j] = bg[j] = 0.f;
                            }
                            uint4 o;
                            __half2* oh = reinterpret_cast<__half2*>(&o);
                            float r8[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                // the reference rounds proj output, gelu(gate) and the product to fp16
                                const float val = lb_round_h(__uint_as_float(vv[g * 8 + j]) + bv[j]);
                                const float gate = lb_round_h(__uint_as_float(vg[g * 8 + j]) + bg[j]);
                                r8[j] = val * lb_round_h(gelu_erf(gate));
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(r8[2 * j], r8[2 * j + 1]);
                            *reinterpret_cast<uint4*>(p.out + row * p.ldo + o_base + jn) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CL == 2) mbar_arrive_cluster(map_to_cta(&tmem_empty[acc], 0));
                else mbar_arrive(&tmem_empty[acc]);
            }
            if (++acc == kAcc) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();          // the peer may still arrive on this CTA's barriers / read its smem
    if (warp == 2) {
        if (CL == 2) tmem_dealloc_2sm(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- host side ------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode(lb_ctx* ctx, EncodeTiledFn* fn) {
    if (!ctx->tmap_encode) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        LB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres));
        LB_REQUIRE(f != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
        ctx->tmap_encode = f;
    }
    *fn = reinterpret_cast<EncodeTiledFn>(ctx->tmap_encode);
    return 0;
}

// 4-D NHWC activation map: dims (C, W, H, B), box (64, tw, th, tb), 128B swizzle, zero OOB fill.
int encode_act_map(lb_ctx* ctx, CUtensorMap* m, const void* base, int64_t ld, int C, int W, int H, int B, int tw,
                   int th, int tb) {
    EncodeTiledFn enc;
    if (int e = get_encode(ctx, &enc)) return e;
    LB_REQUIRE(lb_aligned16(base), "activation base must be 16-byte aligned");
    LB_REQUIRE(ld % 8 == 0 && ld >= C, "activation row stride must be a multiple of 8 elements and >= C");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * W, (cuuint64_t)ld * 2 * W * H};
    cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation C=%d W=%d H=%d B=%d ld=%lld box=%d,%d,%d) failed: %d",
               C, W, H, B, (long long)ld, tw, th, tb, (int)r);
    return 0;
}

int encode_weight_map(lb_ctx* ctx, CUtensorMap* m, const void* base, int64_t ld, int64_t K, int N, int bn) {
    EncodeTiledFn enc;
    if (int e = get_encode(ctx, &enc)) return e;
    LB_REQUIRE(lb_aligned16(base), "weight base must be 16-byte aligned");
    LB_REQUIRE(ld % 8 == 0 && ld >= K, "weight row stride must be a multiple of 8 elements and >= K");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight K=%lld N=%d bn=%d) failed: %d", (long long)K, N, bn,
               (int)r);
    return 0;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

template <int BN, int CL, int CR> int launch_bn_cl(const GemmPlan& plan, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        LB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, CL, CR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Cfg<BN>::smem_bytes(CR ? Cfg<BN>::cr_stages : Cfg<BN>::deep)));
        if (CR)     // ask for the full shared-memory carve-out so that two ~110 KB CTAs fit
            LB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, CL, CR>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                               cudaSharedmemCarveoutMaxShared));
        attr_set = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)plan.grid);
    cfg.blockDim = dim3(kThreadsGemm);
    cfg.dynamicSmemBytes = Cfg<BN>::smem_bytes(plan.p.stages);
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = lb_pdl_enabled() ? 2 : 1;
    LB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CL, CR>, plan.p));
    return 0;
}
template <int BN> int launch_bn(const GemmPlan& plan, cudaStream_t st) {
    if (plan.cluster == 2) return launch_bn_cl<BN, 2, 0>(plan, st);
    return launch_bn_cl<BN, 1, 0>(plan, st);      // (the CR = 1 co-resident variant is no longer instantiated, see below)
}

}  // namespace

int gemm_plan_build(lb_ctx* ctx, const GemmDesc& d, GemmPlan* plan) {
    LB_REQUIRE(ctx && plan, "gemm: null ctx/plan");
    LB_REQUIRE(d.a0 && d.w && d.out, "gemm: null a0/w/out");
    LB_REQUIRE(d.B >= 1 && d.H >= 1 && d.W >= 1 && d.N >= 8, "gemm: bad shape B=%d H=%d W=%d N=%d", d.B, d.H, d.W, d.N);
    LB_REQUIRE(d.taps == 1 || d.taps == 9, "gemm: taps must be 1 or 9");
    LB_REQUIRE(d.a0_c % kBK == 0 && d.a0_c > 0, "gemm: a0 channels (%d) must be a multiple of 64", d.a0_c);
    LB_REQUIRE(d.a1 == nullptr || (d.a1_c % kBK == 0 && d.a1_c > 0), "gemm: a1 channels must be a multiple of 64");
    LB_REQUIRE(d.N % 8 == 0, "gemm: N (%d) must be a multiple of 8", d.N);
    LB_REQUIRE(d.out_ld % 8 == 0 && lb_aligned16(d.out), "gemm: out must be 16B aligned with ld %% 8 == 0");
    LB_REQUIRE(!d.res || (d.res_ld % 8 == 0 && lb_aligned16(d.res)), "gemm: residual alignment");
    LB_REQUIRE(!d.bias || lb_aligned16(d.bias), "gemm: bias alignment");
    LB_REQUIRE(!d.bias2 || (lb_aligned16(d.bias2) && d.bias2_ld % 8 == 0), "gemm: bias2 alignment");
    GemmParams& p = plan->p;
    memset(&p, 0, sizeof(p));
    // --- M tiling: a 128-row tile is a (tw x th x tb) box of pixels
    int tw, th, tb;
    if (d.W >= kBM || (d.H == 1 && d.B == 1)) {   // rows of a plain matrix: ragged tail is zero-filled by TMA
        tw = kBM; th = 1; tb = 1;
    } else {
        LB_REQUIRE(is_pow2(d.W), "gemm: W (%d) < 128 must be a power of two", d.W);
        tw = d.W;
        th = kBM / tw;
        if (th > d.H) {
            LB_REQUIRE(is_pow2(d.H), "gemm: H (%d) must be a power of two when H*W < 128", d.H);
            th = d.H;
        }
        tb = kBM / (tw * th);
    }
    p.tw = tw; p.th = th; p.tb = tb;
    p.W = d.W; p.H = d.H; p.B = d.B;
    p.tiles_x = (int)lb_ceil_div(d.W, tw);
    p.tiles_y = (int)lb_ceil_div(d.H, th);
    const int tiles_b = (int)lb_ceil_div(d.B, tb);
    p.tiles_m = p.tiles_x * p.tiles_y * tiles_b;
    // --- N tiling
    int bn;
    if ((d.mode & 0xff) == 1) {
        bn = (d.mode & LB_GEMM_GEGLU256) ? 256 : 128;      // the weight rows are interleaved per N tile by the caller
        LB_REQUIRE(d.N % bn == 0, "gemm: GEGLU needs N %% %d == 0 (got %d)", bn, d.N);
    } else if (d.N % 256 == 0 && (int64_t)p.tiles_m * (d.N / 256) >= 2 * ctx->sm_count) bn = 256;
    else if (d.N % 160 == 0) bn = 160;
    else if (d.N % 128 == 0) bn = 128;
    else if (d.N <= 64) bn = 64;
    else bn = 128;
    plan->bn = bn;
    p.tiles_n = (int)lb_ceil_div(d.N, bn);
    p.N = d.N;
    p.mode = d.mode & 0xff;
    p.static_w = (d.mode & LB_GEMM_STATIC_W) ? 1 : 0;
    // --- K segments
    int ns = 0, total = 0;
    if (d.taps == 9) {
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx) {
                p.seg_map[ns] = 0; p.seg_dy[ns] = ky - 1; p.seg_dx[ns] = kx - 1; p.seg_kb[ns] = d.a0_c / kBK;
                total += p.seg_kb[ns++];
            }
    } else {
        p.seg_map[ns] = 0; p.seg_dy[ns] = 0; p.seg_dx[ns] = 0; p.seg_kb[ns] = d.a0_c / kBK;
        total += p.seg_kb[ns++];
    }
    if (d.a1) {
        p.seg_map[ns] = 1; p.seg_dy[ns] = 0; p.seg_dx[ns] = 0; p.seg_kb[ns] = d.a1_c / kBK;
        total += p.seg_kb[ns++];
    }
    p.num_segs = ns;
    p.total_kb = total;
    const int64_t Ktot = (int64_t)total * kBK;
    if (int e = encode_act_map(ctx, &p.tmA[0], d.a0, d.a0_ld, d.a0_c, d.W, d.H, d.B, tw, th, tb)) return e;
    if (d.a1) {
        if (int e = encode_act_map(ctx, &p.tmA[1], d.a1, d.a1_ld, d.a1_c, d.W, d.H, d.B, tw, th, tb)) return e;
    } else {
        p.tmA[1] = p.tmA[0];
    }
    if (int e = encode_weight_map(ctx, &p.tmB, d.w, d.w_ld, Ktot, d.N, bn)) return e;
    if (int e = encode_weight_map(ctx, &p.tmBh, d.w, d.w_ld, Ktot, d.N, bn / 2)) return e;
    p.out = static_cast<__half*>(d.out);
    p.ldo = d.out_ld;
    p.bias = static_cast<const __half*>(d.bias);
    p.bias2 = static_cast<const __half*>(d.bias2);
    p.bias2_ld = d.bias2_ld;
    p.res = static_cast<const __half*>(d.res);
    p.ldr = d.res_ld;
    p.err_flag = lb_err_flag(ctx);
    p.debug = getenv("LB_GEMM_DEBUG") ? atoi(getenv("LB_GEMM_DEBUG")) : 0;
    p.relu = (d.mode & LB_GEMM_RELU) ? 1 : 0;
    LB_REQUIRE(!p.relu || p.mode == 0, "gemm: LB_GEMM_RELU needs the linear epilogue");
    if (d.ln_stats) {
        LB_REQUIRE(d.ln_csum && d.ln_bias && d.ln_parts >= 1 && d.ln_parts <= 64, "gemm: LayerNorm fold needs ln_csum, "
                   "ln_bias and 1 <= ln_parts <= 64");
        LB_REQUIRE(d.taps == 1 && !d.a1 && !d.res && !d.bias2 && !d.bias, "gemm: LayerNorm fold applies to a plain linear "
                   "(its bias is part of ln_bias)");
        LB_REQUIRE(lb_aligned16(d.ln_csum) && lb_aligned16(d.ln_bias) && lb_aligned16(d.ln_stats), "gemm: ln_* alignment");
        p.ln_stats = static_cast<const float2*>(d.ln_stats);
        p.ln_parts = d.ln_parts;
        p.ln_csum = static_cast<const float*>(d.ln_csum);
        p.ln_bias = static_cast<const float*>(d.ln_bias);
        p.ln_inv_k = 1.0f / (float)d.a0_c;
        p.ln_eps = d.ln_eps;
    }
    if (d.stats_out) {
        LB_REQUIRE(p.mode == 0, "gemm: stats_out needs the linear epilogue");
        LB_REQUIRE(d.stats_parts == kEpiParts * p.tiles_n, "gemm: stats_parts must be %d * ceil(N / %d) = %d (got %d)",
                   kEpiParts, bn, kEpiParts * p.tiles_n, d.stats_parts);
        p.stats_out = static_cast<float2*>(d.stats_out);
    }
    // CTA pairs (cta_group::2, M = 256) whenever there are at least two M tiles
    // (measured: the pair wins ~3 % on long-K multi-wave problems -- the big convolutions -- and loses up to 15 %
    //  on short-K / single-wave ones, where its cluster launch + sync overhead dominates)
    const char* force = getenv("LB_GEMM_CLUSTER");
    // (r02g: the pair also wins 7-13 % on the huge-M, N <= 256 VAE-decoder convolutions -- 1M x 128 x 1152 308 -> 287 us,
    //  1M x 256 x 2304 941 -> 836 us, 262144 x 256 x 2304 232 -> 209 us: thousands of tiles, the shared weight tile halves
    //  each SM's operand traffic -- even though their K is short)
    const int64_t n_tiles = (int64_t)p.tiles_m * p.tiles_n;
    plan->cluster = (p.tiles_m >= 2 && n_tiles >= 256 && (total >= 40 || (p.tiles_m >= 256 && n_tiles >= 1024))) ? 2 : 1;
    if (force) plan->cluster = (atoi(force) == 2 && p.tiles_m >= 2) ? 2 : 1;
    if (plan->cluster == 2) {
        const int groups = ((p.tiles_m + 1) / 2) * p.tiles_n;
        const int max_groups = ctx->sm_count / 2;
        plan->grid = 2 * (groups < max_groups ? groups : max_groups);
    } else {
        const int tiles = p.tiles_m * p.tiles_n;
        plan->grid = tiles < ctx->sm_count ? tiles : ctx->sm_count;
    }
    // The co-resident variant (CR = 1: 3-stage ring, one accumulator, two CTAs per SM) was measured in r02b and dropped:
    // with only 3 ring stages the TMA round trip (~1900 clk vs 256-320 clk per k-block) is no longer hidden -- conv
    // 2048x1280x11520 45 -> 75 us, to_out 2048x1280x1280 11.6 -> 16.2 us, UNet step 23.5 -> 26.1 ms
    // (profiles/r02b_bench_ops_coresident_on.txt).
    plan->coresident = 0;
    // ring depth: deep for long-K problems (>= 40 k-blocks: the 3x3 convolutions, FF-out), shallow otherwise
    {
        const bool deep = total >= 40 && !getenv("LB_GEMM_SHALLOW");
        const int bnv = plan->bn;
        const int shallow_st = (bnv <= 64) ? 8 : (bnv <= 128) ? 6 : (bnv <= 160) ? 5 : 4;
        const int deep_st = (bnv <= 64) ? 8 : (bnv <= 128) ? 7 : (bnv <= 160) ? 6 : 4;
        p.stages = deep ? deep_st : shallow_st;
    }
    plan->smem_bytes = 0;
    return 0;
}

int gemm_plan_launch(const GemmPlan& plan, cudaStream_t st) {
    switch (plan.bn) {
        case 64: return launch_bn<64>(plan, st);
        case 128: return launch_bn<128>(plan, st);
        case 160: return launch_bn<160>(plan, st);
        case 256: return launch_bn<256>(plan, st);
    }
    lb_set_error("gemm: unsupported N tile %d", plan.bn);
    return 2;
}

extern "C" int lb_gemm_stats_parts(lb_ctx* ctx, const lb_gemm_desc* desc) {
    if (!ctx || !desc) return -1;
    GemmDesc d = *reinterpret_cast<const GemmDesc*>(desc);
    d.stats_out = nullptr;
    GemmPlan plan;
    if (gemm_plan_build(ctx, d, &plan)) return -1;
    return kEpiParts * plan.p.tiles_n;
}

extern "C" int lb_gemm(lb_ctx* ctx, const lb_gemm_desc* desc, void* stream) {
    LB_REQUIRE(ctx && desc, "lb_gemm: null argument");
    static_assert(sizeof(GemmDesc) == sizeof(lb_gemm_desc), "GemmDesc / lb_gemm_desc layout drift");
    GemmPlan plan;
    if (int e = gemm_plan_build(ctx, *reinterpret_cast<const GemmDesc*>(desc), &plan)) return e;
    return gemm_plan_launch(plan, lb_stream(stream));
}
