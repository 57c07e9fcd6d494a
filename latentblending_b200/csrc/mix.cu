// mix.cu -- K1: latent mixing kernels (whole-row slerp, lerp).
//
// Replaces latentblending/utils.py:29-71 (interpolate_spherical), the
// per-branch parental-mix loop blending_engine.py:442-450 and the in-loop
// crossfeed diffusers_holder.py:322-324.  HBM-bound: 2 reads + 1 write per
// element (6 B/elem in fp16).
//
// Fast path (slerp_cluster_kernel): one thread-block CLUSTER per row.  Each CTA
// keeps its slice of both inputs in registers, the three fp64 row reductions
// (|p0|^2, |p1|^2, <p0,p1>) are combined across the cluster through distributed
// shared memory, and the axpby is applied to the registers -- a single pass
// over HBM with 128-bit streaming loads/stores.
// Generic path (any n / alignment): partial-sum kernel + apply kernel through a
// small workspace; deterministic (fixed summation order, no atomics).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 256;
constexpr int kParts = 8;  // partials per row in the generic path
constexpr double kClampEps = 1e-7;  // utils.py:55

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

__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ void from_f(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void from_f(float* p, float v) { *p = v; }

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

// fp64 axpby without FMA contraction (torch: mul, mul, add), then the
// reference's fp64 -> fp32 -> storage-dtype cast chain.
__device__ __forceinline__ float slerp_elem(float a, float b, double s0, double s1) {
    double r = __dadd_rn(__dmul_rn((double)a, s0), __dmul_rn((double)b, s1));
    return __double2float_rn(r);
}

__device__ __forceinline__ void block_reduce3(double& aa, double& bb, double& ab, double* sm /*[3*8]*/) {
    aa = lb_warp_sum(aa);
    bb = lb_warp_sum(bb);
    ab = lb_warp_sum(ab);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        sm[w] = aa;
        sm[8 + w] = bb;
        sm[16 + w] = ab;
    }
    __syncthreads();
    aa = bb = ab = 0.0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) {
        aa += sm[i];
        bb += sm[8 + i];
        ab += sm[16 + i];
    }
}

// ---- fast path: one cluster per row ----------------------------------------------
template <typename T, int CHUNKS>
__global__ void __launch_bounds__(kThreads)
slerp_cluster_kernel(const T* __restrict__ p0, const T* __restrict__ p1, T* __restrict__ out, int64_t n,
                     int64_t stride0, int64_t stride1, int64_t stride_out, double fract,
                     const double* __restrict__ fract_rows) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int VE = Vec<T>::N;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned crank = cluster.block_rank();
    const unsigned csize = cluster.num_blocks();
    const int64_t row = blockIdx.y;
    const T* a_row = p0 + row * stride0;
    const T* b_row = p1 + row * stride1;
    T* o_row = out + row * stride_out;

    __shared__ double red[24];
    __shared__ double cta_sum[3];

    uint4 va[CHUNKS], vb[CHUNKS];
    int64_t off[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        int64_t v = ((int64_t)(crank * CHUNKS + c)) * kThreads + threadIdx.x;
        off[c] = v * VE;
        if (off[c] < n) {
            va[c] = lb_ldg_stream(a_row + off[c]);
            vb[c] = lb_ldg_stream(b_row + off[c]);
        } else {
            va[c] = make_uint4(0, 0, 0, 0);
            vb[c] = make_uint4(0, 0, 0, 0);
        }
    }
    double aa = 0.0, bb = 0.0, ab = 0.0;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        float fa[VE], fb[VE];
        Vec<T>::unpack(va[c], fa);
        Vec<T>::unpack(vb[c], fb);
#pragma unroll
        for (int e = 0; e < VE; ++e) {
            double da = fa[e], db = fb[e];
            aa = fma(da, da, aa);
            bb = fma(db, db, bb);
            ab = fma(da, db, ab);
        }
    }
    block_reduce3(aa, bb, ab, red);
    if (threadIdx.x == 0) {
        cta_sum[0] = aa;
        cta_sum[1] = bb;
        cta_sum[2] = ab;
    }
    cluster.sync();
    double taa = 0.0, tbb = 0.0, tab = 0.0;
    for (unsigned r = 0; r < csize; ++r) {
        const double* remote = cluster.map_shared_rank(cta_sum, r);
        taa += remote[0];
        tbb += remote[1];
        tab += remote[2];
    }
    cluster.sync();  // nobody may exit while a peer still reads its cta_sum
    const double f = fract_rows ? fract_rows[row] : fract;
    double s0, s1;
    slerp_weights(taa, tbb, tab, f, s0, s1);
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        if (off[c] < n) {
            float fa[VE], fb[VE], fo[VE];
            Vec<T>::unpack(va[c], fa);
            Vec<T>::unpack(vb[c], fb);
#pragma unroll
            for (int e = 0; e < VE; ++e) fo[e] = slerp_elem(fa[e], fb[e], s0, s1);
            lb_stg_stream(o_row + off[c], Vec<T>::pack(fo));
        }
    }
}

// ---- generic path ---------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
slerp_partial_kernel(const T* __restrict__ p0, const T* __restrict__ p1, int64_t n, int64_t stride0,
                     int64_t stride1, double* __restrict__ partials) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t row = blockIdx.y;
    const T* a_row = p0 + row * stride0;
    const T* b_row = p1 + row * stride1;
    __shared__ double red[24];
    double aa = 0.0, bb = 0.0, ab = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)kParts * kThreads) {
        double da = to_f(a_row[i]), db = to_f(b_row[i]);
        aa = fma(da, da, aa);
        bb = fma(db, db, bb);
        ab = fma(da, db, ab);
    }
    block_reduce3(aa, bb, ab, red);
    if (threadIdx.x == 0) {
        double* dst = partials + (row * kParts + blockIdx.x) * 3;
        dst[0] = aa;
        dst[1] = bb;
        dst[2] = ab;
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
slerp_apply_kernel(const T* __restrict__ p0, const T* __restrict__ p1, T* __restrict__ out, int64_t n,
                   int64_t stride0, int64_t stride1, int64_t stride_out, double fract,
                   const double* __restrict__ fract_rows, const double* __restrict__ partials) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t row = blockIdx.y;
    double aa = 0.0, bb = 0.0, ab = 0.0;
    for (int p = 0; p < kParts; ++p) {
        const double* src = partials + (row * kParts + p) * 3;
        aa += src[0];
        bb += src[1];
        ab += src[2];
    }
    double s0, s1;
    slerp_weights(aa, bb, ab, fract_rows ? fract_rows[row] : fract, s0, s1);
    const T* a_row = p0 + row * stride0;
    const T* b_row = p1 + row * stride1;
    T* o_row = out + row * stride_out;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        from_f(o_row + i, slerp_elem(to_f(a_row[i]), to_f(b_row[i]), s0, s1));
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
lerp_kernel(const T* __restrict__ p0, const T* __restrict__ p1, T* __restrict__ out, int64_t n, float w0,
            float w1) {
    pdl_launch_dependents();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        // torch: (1-f)*p0 -> store dtype; f*p1 -> store dtype; add -> store dtype
        T a, b;
        from_f(&a, w0 * to_f(p0[i]));
        from_f(&b, w1 * to_f(p1[i]));
        from_f(out + i, __fadd_rn(to_f(a), to_f(b)));
    }
}

template <typename T, int CHUNKS>
int launch_cluster(const T* p0, const T* p1, T* out, int64_t rows, int64_t n, int64_t s0, int64_t s1,
                   int64_t so, double fract, const double* fract_rows, int csize, cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(csize, (unsigned)rows, 1);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    LB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, slerp_cluster_kernel<T, CHUNKS>, p0, p1, out, n, s0, s1, so, fract,
                                     fract_rows));
    return 0;
}

template <typename T>
int slerp_dispatch(const void* p0v, const void* p1v, void* outv, int64_t rows, int64_t n, int64_t s0,
                   int64_t s1, int64_t so, double fract, const double* fract_rows, void* ws, cudaStream_t st) {
    const T* p0 = static_cast<const T*>(p0v);
    const T* p1 = static_cast<const T*>(p1v);
    T* out = static_cast<T*>(outv);
    constexpr int VE = Vec<T>::N;
    const bool vec_ok = (n % VE == 0) && (s0 % VE == 0) && (s1 % VE == 0) && (so % VE == 0) &&
                        lb_aligned16(p0) && lb_aligned16(p1) && lb_aligned16(out);
    const int64_t vecs = n / VE;
    // cluster of up to 8 CTAs x 256 threads x CHUNKS vectors
    if (vec_ok && vecs <= 8 * kThreads * 8 && rows <= 65535) {
        const int64_t per_cta1 = kThreads;  // vectors per CTA at CHUNKS=1
        int csize = 1;
        while (csize < 8 && (int64_t)csize * per_cta1 * 2 < vecs) csize *= 2;  // prefer >=2 chunks/thread before growing
        while (csize < 8 && (int64_t)csize * per_cta1 * 8 < vecs) csize *= 2;
        int64_t need = lb_ceil_div(vecs, (int64_t)csize * per_cta1);
        if (need <= 1) return launch_cluster<T, 1>(p0, p1, out, rows, n, s0, s1, so, fract, fract_rows, csize, st);
        if (need <= 2) return launch_cluster<T, 2>(p0, p1, out, rows, n, s0, s1, so, fract, fract_rows, csize, st);
        if (need <= 4) return launch_cluster<T, 4>(p0, p1, out, rows, n, s0, s1, so, fract, fract_rows, csize, st);
        return launch_cluster<T, 8>(p0, p1, out, rows, n, s0, s1, so, fract, fract_rows, csize, st);
    }
    LB_REQUIRE(ws != nullptr, "lb_slerp_rows: generic path needs the workspace");
    LB_REQUIRE(rows <= 65535, "lb_slerp_rows: rows > 65535 unsupported");
    double* partials = static_cast<double*>(ws);
    lb_launch_pdl(slerp_partial_kernel<T>, dim3(kParts, (unsigned)rows), kThreads, 0, st, p0, p1, n, s0, s1, partials);
    LB_LAUNCH_CHECK();
    unsigned gx = (unsigned)lb_ceil_div(n, (int64_t)kThreads * 8);
    if (gx < 1) gx = 1;
    if (gx > 1024) gx = 1024;
    lb_launch_pdl(slerp_apply_kernel<T>, dim3(gx, (unsigned)rows), kThreads, 0, st, p0, p1, out, n, s0, s1, so, fract,
                                                                         fract_rows, partials);
    LB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" size_t lb_slerp_workspace_bytes(int64_t rows, int64_t /*n*/) {
    return (size_t)rows * kParts * 3 * sizeof(double);
}

extern "C" int lb_slerp_rows(lb_ctx* ctx, const void* p0, const void* p1, void* out, int64_t rows, int64_t n,
                             int64_t stride0, int64_t stride1, int64_t stride_out, int dtype, double fract,
                             const double* fract_rows_dev, void* workspace_dev, void* stream) {
    LB_REQUIRE(ctx != nullptr, "lb_slerp_rows: null context");
    LB_REQUIRE(rows >= 0 && n >= 0, "lb_slerp_rows: negative size");
    if (rows == 0 || n == 0) return 0;
    LB_REQUIRE(p0 && p1 && out, "lb_slerp_rows: null buffer");
    LB_REQUIRE(dtype == 0 || dtype == 1, "lb_slerp_rows: dtype must be 0 (fp16) or 1 (fp32)");
    cudaStream_t st = lb_stream(stream);
    if (dtype == 0)
        return slerp_dispatch<__half>(p0, p1, out, rows, n, stride0, stride1, stride_out, fract, fract_rows_dev,
                                      workspace_dev, st);
    return slerp_dispatch<float>(p0, p1, out, rows, n, stride0, stride1, stride_out, fract, fract_rows_dev,
                                 workspace_dev, st);
}

extern "C" int lb_lerp(lb_ctx* ctx, const void* p0, const void* p1, void* out, int64_t n, int dtype,
                       double fract, void* stream) {
    LB_REQUIRE(ctx != nullptr, "lb_lerp: null context");
    if (n == 0) return 0;
    LB_REQUIRE(p0 && p1 && out, "lb_lerp: null buffer");
    LB_REQUIRE(dtype == 0 || dtype == 1, "lb_lerp: dtype must be 0 (fp16) or 1 (fp32)");
    const float w0 = (float)(1.0 - fract), w1 = (float)fract;
    unsigned grid = (unsigned)lb_ceil_div(n, kThreads * 4);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    cudaStream_t st = lb_stream(stream);
    if (dtype == 0)
        lb_launch_pdl(lerp_kernel<__half>, grid, kThreads, 0, st, (const __half*)p0, (const __half*)p1, (__half*)out, n, w0, w1);
    else
        lb_launch_pdl(lerp_kernel<float>, grid, kThreads, 0, st, (const float*)p0, (const float*)p1, (float*)out, n, w0, w1);
    LB_LAUNCH_CHECK();
    return 0;
}
