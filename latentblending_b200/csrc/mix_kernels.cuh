// mix_kernels.cuh -- device code of K1 (whole-row slerp).  Included by mix.cu (the C ABI) and by
// tools/ubench_mix.cu (stand-alone micro-benchmark of the variants).
//
// Arithmetic contract = latentblending/utils.py:29-71 (interpolate_spherical):
//   p0,p1 -> fp64; aa = sum p0^2, bb = sum p1^2, ab = sum p0*p1 (fp64);
//   dot = clamp(ab / (sqrt(aa) sqrt(bb)), -1+1e-7, 1-1e-7); theta0 = acos(dot);
//   s0 = sin(theta0 - theta0*f)/sin(theta0); s1 = sin(theta0*f)/sin(theta0);
//   out = (storage dtype)(float)(p0*s0 + p1*s1)         [fp64 mul, mul, add, no FMA contraction]
//
// slerp_stage_kernel (fast path): one thread-block cluster per row.  Each CTA bulk-copies its slice of both
// inputs into shared memory ONCE (cp.async.bulk, 16-128 KiB in flight per CTA with no register cost), so HBM
// sees exactly 2 reads + 1 write per element (6 B/elem in fp16).
//   pass 1 (smem): the three fp64 row sums; combined across the cluster through distributed shared memory in a
//                  fixed order (deterministic, identical on every CTA).
//   pass 2 (smem): the axpby.  The reference evaluates it in fp64 and rounds fp64 -> fp32 -> fp16; doing that per
//                  element costs 3 fp64 conversions + 3 fp64 ops and makes the kernel fp64-pipe-bound, not
//                  HBM-bound.  Instead each element is evaluated in fp32 with the weights split hi+lo
//                  (|error| <= 2^-23 * (|a s0| + |b s1|), proven below) and the fp16 rounding is CERTIFIED: if the
//                  fp32 value is farther from every fp16 rounding boundary than the error bound, rounding it gives
//                  bit-for-bit the reference result; otherwise (~0.3 % of elements, subnormal / overflowing
//                  results, NaN/Inf) that element takes the exact fp64 path.  Output is bit-identical to the
//                  all-fp64 evaluation.
#pragma once
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace lbmix {

namespace cg = cooperative_groups;

constexpr double kClampEps = 1e-7;  // utils.py:55

// ---- element packing -------------------------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<__half> {
    static constexpr int N = 8;
    __device__ static void unpack(const uint4& v, float (&f)[8]) {
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 t = __half22float2(h[i]);
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    }
    __device__ static uint4 pack(const float (&f)[8]) {
        uint4 v;
        __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        return v;
    }
};
template <> struct Vec<float> {
    static constexpr int N = 4;
    __device__ static void unpack(const uint4& v, float (&f)[4]) {
        f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
        f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    }
    __device__ static uint4 pack(const float (&f)[4]) {
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                          __float_as_uint(f[3]));
    }
};

// utils.py:54-63 in fp64: the two slerp weights from the three row sums.
__device__ __forceinline__ void slerp_weights(double aa, double bb, double ab, double fract, double& s0,
                                              double& s1) {
    double norm = sqrt(aa) * sqrt(bb);
    double dot = ab / norm;
    dot = fmin(fmax(dot, -1.0 + kClampEps), 1.0 - kClampEps);
    double theta0 = acos(dot);
    double sin0 = sin(theta0);
    double theta_t = theta0 * fract;
    s0 = sin(theta0 - theta_t) / sin0;
    s1 = sin(theta_t) / sin0;
}

// fp64 axpby without FMA contraction (torch: mul, mul, add), then the reference's fp64 -> fp32 cast
// (the storage-dtype cast follows at the call site).
__device__ __forceinline__ float slerp_elem(float a, float b, double s0, double s1) {
    double r = __dadd_rn(__dmul_rn((double)a, s0), __dmul_rn((double)b, s1));
    return __double2float_rn(r);
}

// fp64 weight split into fp32 hi + lo (|s - hi - lo| <= 2^-48 |s|)
struct SplitW {
    float s0h, s0l, s1h, s1l;
    __device__ SplitW(double s0, double s1) {
        s0h = __double2float_rn(s0);
        s0l = __double2float_rn(s0 - (double)s0h);
        s1h = __double2float_rn(s1);
        s1l = __double2float_rn(s1 - (double)s1h);
    }
};

// Certified fp16 result of (half)(float)(fp64(a)*s0 + fp64(b)*s1) for fp16-valued a, b.
//   r = fma(a,s0h, fma(b,s1h, fma(a,s0l, b*s1l))):  with M = |a s0h| + |b s1h| the four roundings contribute
//   2^-48 M, 2^-47 M, 2^-24 M(1+e), 2^-24 M(1+e)  =>  |r - x| <= 2^-23 M (1.01), x the exact real value; the
//   reference's fp64 value R has |R - x| <= 2^-52 M.  E = 2^-22 M is used (2x margin, absorbs M's own rounding).
//   Let u = half the fp16 spacing in r's binade and d = |r - RN16(r)|.  If d + E < u then R lies on the same
//   side of the nearest rounding midpoint as r, and since |R - mid| > 2^-23 M >= ulp32(r)/2 so does RN32(R): the
//   reference chain RN16(RN32(R)) equals RN16(r).  E < u/2 additionally covers R and r straddling a power of two
//   (the finer grid below it).  Outside the normal fp16 range, or for NaN/Inf, the test fails -> exact path.
__device__ __forceinline__ __half slerp_elem_h(float a, float b, const SplitW& w, double s0, double s1) {
    float t = b * w.s1l;
    t = fmaf(a, w.s0l, t);
    t = fmaf(b, w.s1h, t);
    const float r = fmaf(a, w.s0h, t);
    const float M = fmaf(fabsf(a), fabsf(w.s0h), fabsf(b) * fabsf(w.s1h));
    const float E = M * 2.384185791015625e-07f;   // 2^-22
    const __half h = __float2half_rn(r);
    const float d = fabsf(r - __half2float(h));
    const uint32_t rb = __float_as_uint(r);
    const float u = __uint_as_float((rb & 0x7f800000u) - (11u << 23));   // 2^(e-11); garbage if |r| < 2^-14 (rejected below)
    const float ar = fabsf(r);
    const bool ok = (E + fmaxf(d, E) < u) && (ar >= 6.103515625e-05f) && (ar < 65000.0f);
    if (ok) return h;
    return __float2half_rn(slerp_elem(a, b, s0, s1));
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of three doubles; result valid in every thread.  sm: [3 * 32] doubles.
template <int THREADS>
__device__ __forceinline__ void block_reduce3(double& aa, double& bb, double& ab, double* sm) {
    constexpr int W = THREADS / 32;
    aa = warp_sum_d(aa);
    bb = warp_sum_d(bb);
    ab = warp_sum_d(ab);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        sm[w] = aa;
        sm[32 + w] = bb;
        sm[64 + w] = ab;
    }
    __syncthreads();
    aa = bb = ab = 0.0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        aa += sm[i];
        bb += sm[32 + i];
        ab += sm[64 + i];
    }
}

// ---- async bulk copy (TMA 1-D) + mbarrier ----------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred P;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P;\n\t}"
            : "=r"(ok)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- fast path --------------------------------------------------------------------------------
// grid (csize, rows), cluster (csize,1,1).  `slice` = elements per CTA (multiple of Vec<T>::N; the last CTA of a
// row may own fewer).  Dynamic smem: 2 * slice * sizeof(T) (16 B aligned).  EXACT2 = evaluate pass 2 in fp64 for
// every element (reference arithmetic verbatim; used for fp32 storage and as the A/B check of the certified path).
template <typename T, int THREADS, bool EXACT2>
__global__ void __launch_bounds__(THREADS)
slerp_stage_kernel(const T* __restrict__ p0, const T* __restrict__ p1, T* __restrict__ out, int64_t n, int slice,
                   int64_t stride0, int64_t stride1, int64_t stride_out, double fract,
                   const double* __restrict__ fract_rows) {
    griddep_launch_dependents();
    constexpr int VE = Vec<T>::N;
    extern __shared__ __align__(128) uint8_t stage_raw[];
    __shared__ double red[96];
    __shared__ double cta_sum[3];
    __shared__ __align__(8) uint64_t bar;

    cg::cluster_group cluster = cg::this_cluster();
    const unsigned crank = cluster.block_rank();
    const unsigned csize = cluster.num_blocks();
    const int64_t row = blockIdx.y;
    const int64_t e0 = (int64_t)crank * slice;                       // first element of this CTA's slice
    const int mine = (int)max((int64_t)0, min((int64_t)slice, n - e0));   // elements this CTA owns
    const uint32_t bytes = (uint32_t)mine * (uint32_t)sizeof(T);
    T* sa = reinterpret_cast<T*>(stage_raw);
    T* sb = reinterpret_cast<T*>(stage_raw + (size_t)slice * sizeof(T));

    if (threadIdx.x == 0) bar_init(&bar, 1);
    __syncthreads();
    griddep_wait();                                                   // inputs may be the previous kernel's output
    if (threadIdx.x == 0) {
        bar_expect(&bar, 2 * bytes);
        if (bytes) {
            bulk_g2s(sa, p0 + row * stride0 + e0, bytes, &bar);
            bulk_g2s(sb, p1 + row * stride1 + e0, bytes, &bar);
        }
    }
    bar_wait(&bar, 0);

    // ---- pass 1: fp64 row sums (two interleaved accumulator sets shorten the DFMA chains)
    const int nvec = mine / VE;
    double aa0 = 0.0, bb0 = 0.0, ab0 = 0.0, aa1 = 0.0, bb1 = 0.0, ab1 = 0.0;
    for (int v = threadIdx.x; v < nvec; v += THREADS) {
        float fa[VE], fb[VE];
        Vec<T>::unpack(reinterpret_cast<const uint4*>(sa)[v], fa);
        Vec<T>::unpack(reinterpret_cast<const uint4*>(sb)[v], fb);
#pragma unroll
        for (int e = 0; e < VE; e += 2) {
            const double da0 = fa[e], db0 = fb[e], da1 = fa[e + 1], db1 = fb[e + 1];
            aa0 = fma(da0, da0, aa0);
            bb0 = fma(db0, db0, bb0);
            ab0 = fma(da0, db0, ab0);
            aa1 = fma(da1, da1, aa1);
            bb1 = fma(db1, db1, bb1);
            ab1 = fma(da1, db1, ab1);
        }
    }
    double aa = aa0 + aa1, bb = bb0 + bb1, ab = ab0 + ab1;
    block_reduce3<THREADS>(aa, bb, ab, red);
    if (threadIdx.x == 0) {
        cta_sum[0] = aa;
        cta_sum[1] = bb;
        cta_sum[2] = ab;
    }
    cluster_arrive();
    cluster_wait();
    double taa = 0.0, tbb = 0.0, tab = 0.0;
    for (unsigned r = 0; r < csize; ++r) {
        const double* remote = cluster.map_shared_rank(cta_sum, r);
        taa += remote[0];
        tbb += remote[1];
        tab += remote[2];
    }
    cluster_arrive();          // peers may exit once everybody has read their cta_sum (waited for at the end)
    const double f = fract_rows ? fract_rows[row] : fract;
    double s0, s1;
    slerp_weights(taa, tbb, tab, f, s0, s1);
    const SplitW w(s0, s1);

    // ---- pass 2: axpby from smem, 128-bit coalesced stores
    T* o_row = out + row * stride_out + e0;
    for (int v = threadIdx.x; v < nvec; v += THREADS) {
        const uint4 ua = reinterpret_cast<const uint4*>(sa)[v];
        const uint4 ub = reinterpret_cast<const uint4*>(sb)[v];
        float fa[VE], fb[VE];
        Vec<T>::unpack(ua, fa);
        Vec<T>::unpack(ub, fb);
        uint4 o;
        if constexpr (!EXACT2 && sizeof(T) == 2) {
            __half* oh = reinterpret_cast<__half*>(&o);
#pragma unroll
            for (int e = 0; e < VE; ++e) oh[e] = slerp_elem_h(fa[e], fb[e], w, s0, s1);
        } else {
            float fo[VE];
#pragma unroll
            for (int e = 0; e < VE; ++e) fo[e] = slerp_elem(fa[e], fb[e], s0, s1);
            o = Vec<T>::pack(fo);
        }
        stg_stream(o_row + (int64_t)v * VE, o);
    }
    cluster_wait();
}

}  // namespace lbmix
