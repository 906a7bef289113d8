// rdf_device.cuh -- device helpers shared by the pair kernels (rdf.cu, rdf_cells.cu).
// See rdf.cu for the statement of the reference arithmetic these implement.
#pragma once

#include <cuda_pipeline.h>
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace rdfdev {

constexpr int kThreads = 256;          // 8 warps
constexpr int kWarps = kThreads / 32;
constexpr int kIPT = 2;                // i-particles per thread
constexpr int kTile = kThreads * kIPT; // 512: i-tile == j-tile (same-group symmetry)
constexpr double kMagic = 6755399441055744.0;  // 1.5 * 2^52

// ---- pack: float[F][n][3] -> float4[F][npad] (x, y, z, exclusion block id) ------

static __global__ void rdf_pack_kernel(const float *__restrict__ raw, int64_t frame_stride,
                                float4 *__restrict__ out, int64_t n, int64_t npad,
                                int64_t excl, int drop_axis)
{
    const int frame = blockIdx.y;
    const float *src = raw + (int64_t)frame * frame_stride;
    float4 *dst = out + (int64_t)frame * npad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npad;
         i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n) {
            v.x = src[3 * i];
            v.y = src[3 * i + 1];
            v.z = src[3 * i + 2];
            if (drop_axis == 0) v.x = 0.f;
            if (drop_axis == 1) v.y = 0.f;
            if (drop_axis == 2) v.z = 0.f;
            v.w = __int_as_float((int)(excl > 0 ? i / excl : i));
        }
        dst[i] = v;
    }
}

// ---- the reference arithmetic, one coordinate ---------------------------------

__device__ __forceinline__ double min_image_sq(float a, float b, double box, double inv)
{
    const float df = __fsub_rn(b, a);
    const double d = (double)df;
    const double s = __dmul_rn(inv, d);
    const double r = __dsub_rn(__dadd_rn(s, kMagic), kMagic);
    const double m = __dmul_rn(box, __dsub_rn(s, r));
    return __dmul_rn(m, m);
}

__device__ __forceinline__ double pair_d2(float xi, float yi, float zi, const float4 &pj,
                                          const FrameBox &fb)
{
    const double sx = min_image_sq(xi, pj.x, fb.box[0], fb.inv[0]);
    const double sy = min_image_sq(yi, pj.y, fb.box[1], fb.inv[1]);
    const double sz = min_image_sq(zi, pj.z, fb.box[2], fb.inv[2]);
    return __dadd_rn(__dadd_rn(sx, sy), sz);
}

// ---- bin lookup ---------------------------------------------------------------
// sT2[k] = (T[k], T[k+1]), k in [0, n_bins).  Returns k in [0, n_bins) or n_bins
// ("not counted": below T[0], at or above T[n_bins], or NaN).

static __device__ __noinline__ int bin_search(double d2, const double2 *sT2, int n_bins)
{
    if (!(d2 >= sT2[0].x) || !(d2 < sT2[n_bins - 1].y)) return n_bins;
    int lo = 0, hi = n_bins;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (d2 >= sT2[mid].x) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int bin_index(double d2, const double2 *sT2, int n_bins,
                                         float g_scale, float g_off)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__double2float_rn(d2)));
    int k = __float2int_rd(fmaf(r, g_scale, g_off));
    k = min(max(k, 0), n_bins - 1);
    const double2 t = sT2[k];
    if (!(d2 >= t.x && d2 < t.y)) k = bin_search(d2, sT2, n_bins);
    return k;
}

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

template <int HIST>
__host__ __device__ inline size_t pair_smem_bytes(int n_bins, int n_words)
{
    size_t b = align16(sizeof(double2) * n_bins) + 2 * kTile * sizeof(float4);
    if (HIST == MDH_HIST_WARP_ATOMIC) b += sizeof(unsigned) * kWarps * n_bins;
    else b += sizeof(unsigned) * ((size_t)kWarps * n_words * 32 + n_bins);
    return b;
}

// Lane-private packed histograms: lane l of warp w owns the words
// priv[(w*n_words + word)*32 + l]; each word holds four 8-bit counters, so the
// read-modify-write is bank-conflict free and needs no atomics.  A lane makes at
// most 254 increments between flushes, so no counter can overflow.
__device__ __forceinline__ void priv_flush(unsigned *priv_w, unsigned *bhist, int n_words,
                                           int n_bins, int lane, unsigned weight)
{
    for (int w = 0; w < n_words; ++w) {
        const unsigned v = priv_w[w * 32 + lane];
        if (__any_sync(0xffffffffu, v != 0)) {
            priv_w[w * 32 + lane] = 0;
            const unsigned a = __reduce_add_sync(0xffffffffu, v & 0x00ff00ffu);
            const unsigned b = __reduce_add_sync(0xffffffffu, (v >> 8) & 0x00ff00ffu);
            if (lane < 4) {
                unsigned val = (lane & 1) ? b : a;
                val = (lane & 2) ? (val >> 16) : (val & 0xffffu);
                const int bin = 4 * w + lane;
                if (bin < n_bins && val) atomicAdd(&bhist[bin], val * weight);
            }
        }
    }
    __syncwarp();
}


constexpr size_t kMaxSmem = 227 * 1024;

}  // namespace rdfdev
