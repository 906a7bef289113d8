// rdf_tri.cu -- pair histogram for TRICLINIC cells (seam #1 with a non-orthogonal box).
//
// The reference hands the full `dims` (lengths and angles) to capped_distance
// (/root/reference/src/mdhelper/analysis/structure.py:93-96); for a triclinic cell the
// third-party MDAnalysis code then (restated from its published algorithm, SURVEY.md
// Appendix A; NOT pinned against MDAnalysis itself, which is not installable here):
//   1. wraps both coordinate sets into the primary cell (lib/include/calc_distances.h
//      `_triclinic_pbc`).  Restated mathematically: along c, then b, then a,
//      s = floor(x_k / h_kk), r -= s * h_k, in double, rounded back to float32.  For
//      coordinates that already lie in the cell this is the identity, as in MDAnalysis;
//      for the others its float rounding details are not reproduced.
//   2. per pair: dx = (double)(float)(r_j - r_i), then `minimum_image_triclinic`: the
//      shortest of the 27 images dx + ix a + iy b + iz c, ix, iy, iz in {-1, 0, 1}
//      (loop order ix, iy, iz; strict "<"), every sum in double in the order
//      ((dx0 + a_x ix) + b_x iy) + c_x iz, (dx1 + b_y iy) + c_y iz, dx2 + c_z iz,
//      and d^2 = (rx rx + ry ry) + rz rz.
// Binning is the squared-threshold comparison of the orthorhombic kernels (rdf.cu).
//
// This is a correctness path (every pair evaluates 27 images in fp64); it exists so that
// triclinic trajectories are analysed instead of rejected.

#include <algorithm>

#include "rdf_device.cuh"

using namespace rdfdev;

namespace {

struct TriBox {            // per frame: lower-triangular cell matrix, float32 values
    double ax, bx, by, cx, cy, cz;
};

// in place on the packed float4 array
__global__ void tri_wrap_kernel(float4 *__restrict__ p, int64_t npad, int n,
                                const TriBox *__restrict__ boxes)
{
    const int frame = blockIdx.y;
    const TriBox h = boxes[frame];
    float4 *pf = p + (int64_t)frame * npad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 v = pf[i];
        double x = v.x, y = v.y, z = v.z;
        double s = floor(z / h.cz);
        if (s != 0.0) { x -= s * h.cx; y -= s * h.cy; z -= s * h.cz; }
        s = floor(y / h.by);
        if (s != 0.0) { x -= s * h.bx; y -= s * h.by; }
        s = floor(x / h.ax);
        if (s != 0.0) x -= s * h.ax;
        v.x = (float)x; v.y = (float)y; v.z = (float)z;
        pf[i] = v;
    }
}

__device__ __forceinline__ double tri_d2(float xi, float yi, float zi, const float4 &pj,
                                         const TriBox &h)
{
    const double dx0 = (double)__fsub_rn(pj.x, xi);
    const double dx1 = (double)__fsub_rn(pj.y, yi);
    const double dx2 = (double)__fsub_rn(pj.z, zi);
    double best = (double)FLT_MAX;
#pragma unroll
    for (int ix = -1; ix <= 1; ++ix) {
        const double rx = __dadd_rn(dx0, h.ax * ix);            // products with -1, 0, 1: exact
#pragma unroll
        for (int iy = -1; iy <= 1; ++iy) {
            const double ry0 = __dadd_rn(rx, h.bx * iy);
            const double ry1 = __dadd_rn(dx1, h.by * iy);
#pragma unroll
            for (int iz = -1; iz <= 1; ++iz) {
                const double rz0 = __dadd_rn(ry0, h.cx * iz);
                const double rz1 = __dadd_rn(ry1, h.cy * iz);
                const double rz2 = __dadd_rn(dx2, h.cz * iz);
                const double dsq = __dadd_rn(__dadd_rn(__dmul_rn(rz0, rz0), __dmul_rn(rz1, rz1)),
                                             __dmul_rn(rz2, rz2));
                if (dsq < best) best = dsq;
            }
        }
    }
    return best;
}

struct TriParams {
    const float4 *p1, *p2;
    int64_t pad1, pad2;
    int n1, n2;
    const TriBox *boxes;
    const double *thr;
    int n_bins;
    unsigned long long *counts;
    int jchunk;                    // j particles per block (multiple of 256)
};

template <bool EXCL>
__global__ void __launch_bounds__(256) rdf_triclinic_kernel(const TriParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double *sT = reinterpret_cast<double *>(smem);
    float4 *sJ = reinterpret_cast<float4 *>(smem + align16(sizeof(double) * (P.n_bins + 1)));
    unsigned *sH = reinterpret_cast<unsigned *>(sJ + 256);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int frame = blockIdx.z;
    const int n_bins = P.n_bins;
    for (int k = tid; k <= n_bins; k += 256) sT[k] = P.thr[k];
    for (int k = tid; k < kWarps * n_bins; k += 256) sH[k] = 0;

    const float4 *f1 = P.p1 + (int64_t)frame * P.pad1;
    const float4 *f2 = P.p2 + (int64_t)frame * P.pad2;
    const TriBox h = P.boxes[frame];
    const int i = blockIdx.x * 256 + tid;
    const bool valid = i < P.n1;
    const float4 a = f1[min(i, P.n1 - 1)];
    unsigned *myhist = sH + warp * n_bins;

    const int j0 = blockIdx.y * P.jchunk, j1 = min(P.n2, j0 + P.jchunk);
    for (int jt = j0; jt < j1; jt += 256) {
        __syncthreads();
        if (jt + tid < j1) sJ[tid] = f2[jt + tid];
        __syncthreads();
        const int jn = min(256, j1 - jt);
        if (!valid) continue;
        for (int jj = 0; jj < jn; ++jj) {
            const float4 pj = sJ[jj];
            if (EXCL && __float_as_int(a.w) == __float_as_int(pj.w)) continue;
            const double d2 = tri_d2(a.x, a.y, a.z, pj, h);
            const int slot = slot_search(d2, sT, n_bins);
            if ((unsigned)(slot - 1) < (unsigned)n_bins) atomicAdd(&myhist[slot - 1], 1u);
        }
    }
    __syncthreads();
    for (int k = tid; k < n_bins; k += 256) {
        unsigned long long s = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += sH[w * n_bins + k];
        if (s) atomicAdd(&P.counts[k], s);
    }
}

}  // namespace

// pack kernel of rdf_device.cuh is instantiated in this translation unit as well
int rdf_accumulate_triclinic_impl(mdh_ctx *c, const float *pos1, int64_t s1, const float *pos2,
                                  int64_t s2, int location, const float *box9, int n_frames)
{
    RdfState &R = c->rdf;
    MDH_REQUIRE(R.configured, MDH_ESTATE, "rdf: accumulate before configure");
    MDH_REQUIRE(n_frames >= 1 && n_frames <= 65535, MDH_EINVAL,
                "rdf: n_frames per call must be in [1, 65535]");
    MDH_REQUIRE(box9 != nullptr, MDH_EINVAL, "rdf: box matrix is NULL");
    MDH_REQUIRE(location == MDH_HOST || location == MDH_DEVICE, MDH_EINVAL,
                "rdf: invalid location");
    MDH_REQUIRE(pos1 != nullptr && (R.same || pos2 != nullptr), MDH_EINVAL,
                "rdf: coordinate pointer is NULL");
    MDH_REQUIRE(s1 >= 3 * R.n1 && (R.same || s2 >= 3 * R.n2), MDH_EINVAL,
                "rdf: frame_stride < 3*n");
    MDH_REQUIRE(R.drop_axis < 0, MDH_EINVAL, "rdf: drop_axis needs an orthorhombic cell");
    const size_t smem = align16(sizeof(double) * (R.n_bins + 1)) + 256 * sizeof(float4) +
                        sizeof(unsigned) * kWarps * (size_t)R.n_bins;
    MDH_REQUIRE(smem <= kMaxSmem, MDH_EINVAL, "rdf: too many bins for the triclinic kernel");

    std::vector<TriBox> hb(n_frames);
    for (int f = 0; f < n_frames; ++f) {
        const float *m = box9 + 9 * f;
        MDH_REQUIRE(m[1] == 0.f && m[2] == 0.f && m[5] == 0.f, MDH_EINVAL,
                    "rdf: the cell matrix of frame %d is not lower triangular", f);
        MDH_REQUIRE(m[0] > 0.f && m[4] > 0.f && m[8] > 0.f && std::isfinite(m[0]) &&
                    std::isfinite(m[4]) && std::isfinite(m[8]), MDH_EINVAL,
                    "rdf: the cell matrix of frame %d is not a valid cell", f);
        hb[f] = TriBox{(double)m[0], (double)m[3], (double)m[4], (double)m[6], (double)m[7],
                       (double)m[8]};
    }
    DevBuf &d_box = R.cell[0];
    if (int rc = d_box.reserve(sizeof(TriBox) * n_frames)) return rc;
    MDH_CUDA(cudaMemcpyAsync(d_box.p, hb.data(), sizeof(TriBox) * n_frames,
                             cudaMemcpyHostToDevice, c->stream));
    MDH_CUDA(cudaStreamSynchronize(c->stream));        // hb is a local

    const int64_t pad1 = (R.n1 + 255) / 256 * 256, pad2 = (R.n2 + 255) / 256 * 256;
    const int n_groups = R.same ? 1 : 2;
    for (int g = 0; g < n_groups; ++g) {
        const float *pos = g ? pos2 : pos1;
        const int64_t stride = g ? s2 : s1, n = g ? R.n2 : R.n1, npad = g ? pad2 : pad1;
        DevBuf &raw = g ? R.raw2[0] : R.raw1[0];
        DevBuf &pk = g ? R.pk2 : R.pk1;
        const float *dsrc = pos;
        int64_t dstride = stride;
        if (location == MDH_HOST) {
            if (int rc = raw.reserve(sizeof(float) * 3 * n * n_frames)) return rc;
            MDH_CUDA(mdh_copy_frames(raw.p, sizeof(float) * 3 * n, pos, sizeof(float) * stride,
                                       sizeof(float) * 3 * n, n_frames, cudaMemcpyHostToDevice,
                                       c->stream));
            dsrc = raw.as<float>();
            dstride = 3 * n;
        }
        if (int rc = pk.reserve(sizeof(float4) * npad * n_frames)) return rc;
        dim3 grid((unsigned)std::min<int64_t>((npad + 255) / 256, 2048), n_frames);
        rdf_pack_kernel<<<grid, 256, 0, c->stream>>>(dsrc, dstride, pk.as<float4>(), n, npad,
                                                     g ? R.excl2 : R.excl1, -1, nullptr, nullptr);
        MDH_CUDA(cudaGetLastError());
        tri_wrap_kernel<<<grid, 256, 0, c->stream>>>(pk.as<float4>(), npad, (int)n,
                                                     d_box.as<TriBox>());
        MDH_CUDA(cudaGetLastError());
        c->launches += 2;
    }

    TriParams P;
    P.p1 = R.pk1.as<float4>();
    P.p2 = R.same ? P.p1 : R.pk2.as<float4>();
    P.pad1 = pad1; P.pad2 = R.same ? pad1 : pad2;
    P.n1 = (int)R.n1; P.n2 = (int)R.n2;
    P.boxes = d_box.as<TriBox>();
    P.thr = R.thr.as<double>();
    P.n_bins = R.n_bins;
    P.counts = R.counts.as<unsigned long long>();
    // enough blocks to fill the device; a block's u32 words cannot overflow (256 x jchunk)
    const int64_t iblocks = (R.n1 + 255) / 256;
    int64_t jsplit = std::max<int64_t>(1, (4 * c->sm_count + iblocks * n_frames - 1) /
                                              (iblocks * n_frames));
    int64_t jchunk = ((R.n2 + jsplit - 1) / jsplit + 255) / 256 * 256;
    jchunk = std::min<int64_t>(jchunk, 1 << 22);
    P.jchunk = (int)jchunk;
    dim3 grid((unsigned)iblocks, (unsigned)((R.n2 + jchunk - 1) / jchunk), (unsigned)n_frames);
    MDH_REQUIRE(grid.y <= 65535, MDH_EINVAL, "rdf: group 2 is too large for the triclinic kernel");
    if (int rc = c->t_rdf.begin(c->stream)) return rc;
    if (R.excl1 > 0) {
        MDH_CUDA(cudaFuncSetAttribute(rdf_triclinic_kernel<true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rdf_triclinic_kernel<true><<<grid, 256, smem, c->stream>>>(P);
    } else {
        MDH_CUDA(cudaFuncSetAttribute(rdf_triclinic_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rdf_triclinic_kernel<false><<<grid, 256, smem, c->stream>>>(P);
    }
    MDH_CUDA(cudaGetLastError());
    c->launches++;
    R.evals += R.n1 * R.n2 * (int64_t)n_frames;
    return c->t_rdf.end(c->stream);
}
