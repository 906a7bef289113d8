// rdf_cells.cu -- cell-list variant of the pair histogram (cut-off runs).
//
// Same per-pair arithmetic and binning as rdf.cu (rdf_device.cuh); only the set
// of candidate pairs shrinks: each frame's particles are counting-sorted into
// cells of edge >= r_cut*(1+1e-5) and every i visits the 27 surrounding cells
// of the j group.  This is the role MDAnalysis' grid search ("nsgrid") plays
// behind capped_distance for the reference (call site
// /root/reference/src/mdhelper/analysis/structure.py:93-96; SURVEY.md Appendix A
// items 2 and 4).  Counts are identical to the all-pairs kernel by construction:
// a pair closer than r_cut always lies in adjacent cells, and every other
// candidate falls above the last threshold and is not counted.
//
// Ordered pairs are enumerated directly (i over group 1, j over group 2, both
// orders and the self pair when the groups coincide), as the reference counts them.

#include <algorithm>

#include "rdf_device.cuh"

using namespace rdfdev;

namespace {

struct CellGrid {          // per frame
    double box[3];
    double inv_w[3];       // nc / box
    int nc[3];
    int ncell;
};

__device__ __forceinline__ int cell_coord(float x, double box, double inv_w, int nc)
{
    double w = (double)x;
    w -= floor(w / box) * box;               // into [0, box) for cell assignment only
    int c = (int)(w * inv_w);
    return min(max(c, 0), nc - 1);
}

__device__ __forceinline__ int cell_id(const float4 &p, const CellGrid &g, int &cx, int &cy,
                                       int &cz)
{
    cx = cell_coord(p.x, g.box[0], g.inv_w[0], g.nc[0]);
    cy = cell_coord(p.y, g.box[1], g.inv_w[1], g.nc[1]);
    cz = cell_coord(p.z, g.box[2], g.inv_w[2], g.nc[2]);
    return (cz * g.nc[1] + cy) * g.nc[0] + cx;
}

// count particles per cell; the old counter value is the particle's rank in its cell
__global__ void cells_count_kernel(const float4 *__restrict__ p, int64_t npad, int n,
                                   const CellGrid *__restrict__ grids, int *__restrict__ cnt,
                                   int cstride, int *__restrict__ rank)
{
    const int frame = blockIdx.y;
    const CellGrid g = grids[frame];
    const float4 *pf = p + (int64_t)frame * npad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int cx, cy, cz;
        const int c = cell_id(pf[i], g, cx, cy, cz);
        rank[(int64_t)frame * n + i] = atomicAdd(&cnt[(int64_t)frame * cstride + c], 1);
    }
}

// exclusive scan of the per-cell counts, one block per frame: tiles of 4096 counters
// (one int4 per thread, coalesced), block scan by warp shuffles, running carry
__global__ void __launch_bounds__(1024) cells_scan_kernel(const int *__restrict__ cnt,
                                                          int *__restrict__ start, int cstride,
                                                          const CellGrid *__restrict__ grids)
{
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    const int frame = blockIdx.x;
    const int ncell = grids[frame].ncell;
    const int *c = cnt + (int64_t)frame * cstride;
    int *s = start + (int64_t)frame * cstride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < ncell; base += 4096) {
        const int k0 = base + 4 * tid;
        int v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = k0 + j < ncell ? c[k0 + j] : 0;
        const int local = v[0] + v[1] + v[2] + v[3];
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int carry = carry_s;
        int run = carry + incl - local + (warp ? warp_sums[warp - 1] : 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (k0 + j < ncell) s[k0 + j] = run;
            run += v[j];
        }
        __syncthreads();
        if (tid == 1023) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
    if (tid == 0) s[ncell] = carry_s;
}

// sorted: float4[F][n] cell-sorted particles.  pairs (optional): the same order in the
// layout the fp32-filter kernel reads, float4[F][2 * npair] with entry 2q = (x0, x1, y0,
// y1) and entry 2q + 1 = (z0, z1, id0, id1) of the sorted particles 2q and 2q + 1, so
// that one 16-byte load yields operands already packed for f32x2 arithmetic.
__global__ void cells_scatter_kernel(const float4 *__restrict__ p, int64_t npad, int n,
                                     const CellGrid *__restrict__ grids,
                                     const int *__restrict__ start, int cstride,
                                     const int *__restrict__ rank, float4 *__restrict__ sorted,
                                     float *__restrict__ pairs, int npair)
{
    const int frame = blockIdx.y;
    const CellGrid g = grids[frame];
    const float4 *pf = p + (int64_t)frame * npad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 v = pf[i];
        int cx, cy, cz;
        const int c = cell_id(v, g, cx, cy, cz);
        const int dst = start[(int64_t)frame * cstride + c] + rank[(int64_t)frame * n + i];
        sorted[(int64_t)frame * n + dst] = v;
        if (pairs) {
            float *q = pairs + ((int64_t)frame * npair + (dst >> 1)) * 8 + (dst & 1);
            q[0] = v.x; q[2] = v.y; q[4] = v.z; q[6] = v.w;
            if (dst == n - 1 && (n & 1)) {       // odd count: finite filler in the last slot
                q[1] = v.x; q[3] = v.y; q[5] = v.z; q[7] = v.w;
            }
        }
    }
}

struct CellParams {
    const float4 *s1, *s2;        // cell-sorted particles, [F][n]
    int n1, n2;
    const int *start2;            // [F][cstride]
    int cstride;
    const CellGrid *grids;
    const FrameBox *boxes;
    const double *thr;
    int n_bins;
    BinGuess guess;
    unsigned long long *counts;
    unsigned long long *evals;
    int half;                     // same group: half stencil, weight 2
    // fp32 filter (rdf_cells_filter_kernel); filt == nullptr: exact kernel does it all
    const float4 *pairs2;         // pair-interleaved copy of s2, [F][2 * npair2]
    int npair2;
    const FrameFilter *filt;
    FilterConst fc;
    int fast_bins;
    unsigned long long *fstats;
};

template <int HIST>
__host__ __device__ inline size_t cells_smem_bytes(int n_bins)
{
    return align16(sizeof(double) * (n_bins + 1)) + hist_smem_bytes<HIST>(n_bins);
}

template <int HIST, bool EXCL, bool FAST>
__global__ void __launch_bounds__(kThreads, 2) rdf_cells_kernel(const CellParams P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double *sT = reinterpret_cast<double *>(smem);
    unsigned *sH =
        reinterpret_cast<unsigned *>(smem + align16(sizeof(double) * (P.n_bins + 1)));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.y;
    // frames the fp32-filter kernel has taken are not done again
    if (P.filt != nullptr && P.filt[frame].wlim != 0u) return;
    const int n_bins = P.n_bins;
    const int n_words = priv_words(n_bins);
    for (int k = tid; k <= n_bins; k += kThreads) sT[k] = P.thr[k];
    const int n_hist_words = (int)(hist_smem_bytes<HIST>(n_bins) / sizeof(unsigned));
    for (int k = tid; k < n_hist_words; k += kThreads) sH[k] = 0;
    __syncthreads();

    unsigned *myhist = (HIST == MDH_HIST_WARP_ATOMIC)
                           ? sH + warp * warp_hist_words(n_bins) + 1
                           : sH + (size_t)warp * n_words * 32;
    unsigned *bhist = sH + (size_t)kWarps * n_words * 32;
    unsigned char *lane_base = reinterpret_cast<unsigned char *>(myhist) + 4 * lane;
    const unsigned hist32 = (unsigned)__cvta_generic_to_shared(myhist);
    const unsigned trash32 = hist32 + 4u * (unsigned)(n_bins + lane);
    const BinGuess guess = P.guess;

    const CellGrid g = P.grids[frame];
    const FrameBox fb = P.boxes[frame];
    const float4 *s1 = P.s1 + (int64_t)frame * P.n1;
    const float4 *s2 = P.s2 + (int64_t)frame * P.n2;
    const int *start = P.start2 + (int64_t)frame * P.cstride;

    const int i = blockIdx.x * kThreads + tid;
    const bool valid = i < P.n1;
    const float4 pi = s1[min(i, P.n1 - 1)];
    const int gi = __float_as_int(pi.w);
    int cx, cy, cz;
    cell_id(pi, g, cx, cy, cz);

    int steps = 0;                       // warp-uniform: increments since the last flush
    unsigned long long my_evals = 0;

    // One range of candidate partners [b, b + len) per lane; the loop bound is the
    // warp maximum so that flushes stay warp-uniform.  The first partner carries
    // weight w_first, the others w_rest (half-stencil runs: 1 for the self pair, 2
    // for every unordered pair).  The next partner is fetched one iteration ahead.
    auto sweep = [&](int b, int len, unsigned w_first, unsigned w_rest) {
        const int maxlen = __reduce_max_sync(0xffffffffu, len);
        my_evals += len;
        float4 nxt = __ldg(s2 + (len > 0 ? b : 0));
        for (int t = 0; t < maxlen; ++t) {
            const bool act = t < len;
            const float4 pj = nxt;
            nxt = __ldg(s2 + (t + 1 < len ? b + t + 1 : 0));
            const unsigned w = t == 0 ? w_first : w_rest;
            const double d2 = pair_d2(pi.x, pi.y, pi.z, pj, fb);
            const bool keep = act && !(EXCL && gi == __float_as_int(pj.w));
            if (HIST == MDH_HIST_WARP_ATOMIC && FAST) {
                bool below;
                const int j = slot_fast_parts(d2, sT, n_bins, guess, below);
                unsigned a = (below ? hist32 - 4u : hist32) + 4u * (unsigned)j;
                if ((!below && j == n_bins) || !keep) a = trash32;
                red_shared(a, w);
            } else {
                int slot = slot_of<FAST>(d2, sT, n_bins, guess);
                if (!keep) slot = 0;
                if (HIST == MDH_HIST_WARP_ATOMIC) {
                    if ((unsigned)(slot - 1) < (unsigned)n_bins)
                        atomicAdd(&myhist[slot - 1], w);
                } else {
                    priv_add(lane_base, slot, w);
                    steps += 2;
                    if (steps >= 253) {
                        priv_flush(myhist, bhist, n_words, n_bins, lane, 1u);
                        steps = 0;
                    }
                }
            }
        }
    };
    auto wrap = [](int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); };
    // the three x-adjacent cells of row (y, z) are one contiguous range of the sorted
    // array, plus one more cell when the x stencil wraps around the box
    auto sweep_row = [&](int y, int z, unsigned w) {
        const int row = (z * g.nc[1] + y) * g.nc[0];
        const int xa = max(cx - 1, 0), xb = min(cx + 1, g.nc[0] - 1);
        const int b0 = start[row + xa];
        sweep(b0, valid ? start[row + xb + 1] - b0 : 0, w, w);
        const int xw = (cx == 0) ? g.nc[0] - 1 : (cx == g.nc[0] - 1 ? 0 : -1);
        const int bw = xw >= 0 ? start[row + xw] : 0;
        sweep(bw, (valid && xw >= 0) ? start[row + xw + 1] - bw : 0, w, w);
    };

    if (!P.half) {
        // full stencil: every ordered (i, j) pair once
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                sweep_row(wrap(cy + dy, g.nc[1]), wrap(cz + dz, g.nc[2]), 1u);
    } else {
        // same group: half stencil.  Forward offsets (dz, dy, dx) > (0, 0, 0) in
        // lexicographic order visit every unordered pair of distinct cells exactly
        // once (the reverse offset belongs to the partner cell); weight 2 stands for
        // both orders.  The own cell contributes j >= i: the self pair once.
        for (int dy = -1; dy <= 1; ++dy)
            sweep_row(wrap(cy + dy, g.nc[1]), wrap(cz + 1, g.nc[2]), 2u);
        sweep_row(wrap(cy + 1, g.nc[1]), cz, 2u);
        const int row = (cz * g.nc[1] + cy) * g.nc[0];
        const int xr = wrap(cx + 1, g.nc[0]);
        const int br = start[row + xr];
        sweep(br, valid ? start[row + xr + 1] - br : 0, 2u, 2u);
        sweep(i, valid ? start[row + cx + 1] - i : 0, 1u, 2u);
    }
    if (HIST == MDH_HIST_LANE_PRIVATE)
        priv_flush(myhist, bhist, n_words, n_bins, lane, 1u);
    __syncthreads();

    for (int k = tid; k < n_bins; k += kThreads) {
        unsigned long long s = 0;
        if (HIST == MDH_HIST_WARP_ATOMIC) {
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += sH[w * warp_hist_words(n_bins) + 1 + k];
        } else {
            s = bhist[k];
        }
        if (s) atomicAdd(&P.counts[k], s);
    }
    // evaluations actually performed (for the roofline bookkeeping)
    for (int o = 16; o; o >>= 1) my_evals += __shfl_xor_sync(0xffffffffu, my_evals, o);
    if (lane == 0 && my_evals) atomicAdd(P.evals, my_evals);
}


// ---- fp32 filter in front of the exact arithmetic (see rdf_filter.cu for the scheme and
// its error bound): same traversal as rdf_cells_kernel, two neighbours per step with
// packed f32x2 arithmetic; uncertain pairs go to a per-block list that is re-evaluated
// with the fp64 arithmetic when the block has finished its sweeps. -------------------

constexpr int kCellListCap = 1024;
// 128-thread blocks, four per SM: a block ends with its slowest warp, and with the same
// registers per SM smaller blocks lose less to that (barrier stalls were 13 % of the
// samples with 256 threads)
constexpr int kCfThreads = 128;
constexpr int kCfWarps = kCfThreads / 32;

__host__ __device__ inline size_t cells_filter_smem_bytes(int n_bins, int sb)
{
    return align16(sizeof(double) * (n_bins + 1)) +
           sizeof(unsigned) * ((size_t)kCfWarps * (((size_t)(n_bins + 2) << sb) + 32)) +
           sizeof(unsigned) * (kCellListCap + 4);
}

// entry = thread << 24 | valid bits << 22 | pair index q
template <bool EXCL, bool LOWER>
__device__ __noinline__ void cells_filter_fix(const CellParams &P, int frame, unsigned entry,
                                              const double *sT, unsigned hist32, unsigned weight)
{
    const FrameFilter ff = P.filt[frame];
    const FrameBox fb = P.boxes[frame];
    const FilterConst fc = P.fc;
    const int i = blockIdx.x * kCfThreads + (int)(entry >> 24);
    const unsigned vbits = (entry >> 22) & 3u;
    const int q = (int)(entry & 0x3fffffu);
    const float4 pi = P.s1[(int64_t)frame * P.n1 + i];
    const float4 *pp = P.pairs2 + ((int64_t)frame * P.npair2 + q) * 2;
    const float4 A = pp[0], B = pp[1];
    unsigned uu[2];
    filter_eval2<LOWER>(pk2(-pi.x, -pi.x), pk2(-pi.y, -pi.y), pk2(-pi.z, -pi.z), pk2(A.x, A.y),
                        pk2(A.z, A.w), pk2(B.x, B.y), ff, fc.scale, ff.offm, fc.cbits, uu[0],
                        uu[1]);
    const unsigned span_l = LOWER ? fc.span : fc.cbits + fc.span;
    for (int h = 0; h < 2; ++h) {
        if (!((vbits >> h) & 1u)) continue;
        const unsigned u = uu[h];
        if (!(u < span_l)) continue;
        if (!((u & ((1u << fc.k) - 1u)) < ff.wlim)) continue;
        const float4 pj = h ? make_float4(A.y, A.w, B.y, B.w) : make_float4(A.x, A.z, B.x, B.z);
        if (EXCL && __float_as_int(pi.w) == __float_as_int(pj.w)) continue;
        const unsigned word = (LOWER ? u : u - fc.cbits) >> (fc.k - fc.sb);
        const double d2 = pair_d2(pi.x, pi.y, pi.z, pj, fb);
        const int slot = P.fast_bins ? slot_fast(d2, sT, P.n_bins, P.guess)
                                     : slot_search(d2, sT, P.n_bins);
        if ((word >> fc.sb) == (unsigned)slot) continue;
        red_shared(hist32 + 4u * word, 0u - weight);
        red_shared(hist32 + 4u * ((unsigned)slot << fc.sb), weight);
    }
}

template <bool EXCL, bool LOWER, bool AUDIT>
__global__ void __launch_bounds__(kCfThreads, 4)
    rdf_cells_filter_kernel(const __grid_constant__ CellParams P)
{
    const int frame = blockIdx.y;
    const FrameFilter ff = P.filt[frame];
    if (ff.wlim == 0u) {                  // left to the exact kernel (whole block)
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&P.fstats[4], 1ull);
        return;
    }
    extern __shared__ __align__(16) unsigned char smem[];
    const int n_bins = P.n_bins;
    const FilterConst fc = P.fc;
    const int hwords = ((n_bins + 2) << fc.sb) + 32;
    double *sT = reinterpret_cast<double *>(smem);
    unsigned *sH = reinterpret_cast<unsigned *>(smem + align16(sizeof(double) * (n_bins + 1)));
    unsigned *sList = sH + kCfWarps * hwords;
    unsigned *sCount = sList + kCellListCap;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k <= n_bins; k += kCfThreads) sT[k] = P.thr[k];
    for (int k = tid; k < kCfWarps * hwords; k += kCfThreads) sH[k] = 0;
    if (tid == 0) *sCount = 0;
    __syncthreads();

    const unsigned hist32 = (unsigned)__cvta_generic_to_shared(sH + warp * hwords);
    const int shift = fc.k - fc.sb;
    const unsigned hbase = hist32 - (LOWER ? 0u : ((fc.cbits >> shift) << 2));
    const unsigned span_l = LOWER ? fc.span : fc.cbits + fc.span;
    const unsigned trash_w = (span_l >> shift) + (unsigned)lane;
    const unsigned fmask = (1u << fc.k) - 1u;
    const float scale = fc.scale, offm = ff.offm;
    const unsigned weight = P.half ? 2u : 1u;     // every sweep of a run has one weight

    const CellGrid g = P.grids[frame];
    const float4 *s1 = P.s1 + (int64_t)frame * P.n1;
    const float4 *pairs = P.pairs2 + (int64_t)frame * P.npair2 * 2;
    const int *start = P.start2 + (int64_t)frame * P.cstride;

    const int i = blockIdx.x * kCfThreads + tid;
    const bool valid = i < P.n1;
    const float4 pi = s1[min(i, P.n1 - 1)];
    const int gi = __float_as_int(pi.w);
    const f32x2 ax = pk2(-pi.x, -pi.x), ay = pk2(-pi.y, -pi.y), az = pk2(-pi.z, -pi.z);
    int cx, cy, cz;
    cell_id(pi, g, cx, cy, cz);

    unsigned long long my_evals = 0, audit_bad = 0, audit_unc = 0;
    const int last_pair = P.npair2 - 1;

    // partners [b, b + len) of this lane, two per step (pair q holds sorted particles 2q
    // and 2q + 1; the ones outside the range get weight 0); warp-uniform trip count
    auto sweep = [&](int b, int len) {
        const int q0 = b >> 1;
        const int nq = len > 0 ? ((b + len + 1) >> 1) - q0 : 0;
        const int maxq = __reduce_max_sync(0xffffffffu, nq);
        my_evals += len;
        const float4 *pp = pairs + 2 * (int64_t)min(q0, last_pair);
        float4 nA = __ldg(pp), nB = __ldg(pp + 1);
        for (int t = 0; t < maxq; ++t) {
            const float4 A = nA, B = nB;
            const int q = q0 + t;
            pp = pairs + 2 * (int64_t)min(q + 1, last_pair);
            nA = __ldg(pp); nB = __ldg(pp + 1);
            // in range?  (lanes past their own nq fall out here as well)
            const bool v0 = (unsigned)(2 * q - b) < (unsigned)len;
            const bool v1 = (unsigned)(2 * q + 1 - b) < (unsigned)len;
            unsigned u0, u1;
            filter_eval2<LOWER>(ax, ay, az, pk2(A.x, A.y), pk2(A.z, A.w), pk2(B.x, B.y), ff, scale,
                                offm, fc.cbits, u0, u1);
            unsigned w0 = min(u0 >> shift, trash_w), w1 = min(u1 >> shift, trash_w);
            if (EXCL && gi == __float_as_int(B.z)) w0 = trash_w;
            if (EXCL && gi == __float_as_int(B.w)) w1 = trash_w;
            red_shared_hot(hbase + (w0 << 2), v0 ? weight : 0u);
            red_shared_hot(hbase + (w1 << 2), v1 ? weight : 0u);
            if (AUDIT) {
                const unsigned uu[2] = {u0, u1};
                const bool vv[2] = {v0, v1};
                for (int h = 0; h < 2; ++h) {
                    if (!vv[h]) continue;
                    const float4 pj = h ? make_float4(A.y, A.w, B.y, B.w)
                                        : make_float4(A.x, A.z, B.x, B.z);
                    const unsigned u = uu[h];
                    const bool unc = (u & fmask) < ff.wlim, in = u < span_l;
                    const double d2 = pair_d2(pi.x, pi.y, pi.z, pj, P.boxes[frame]);
                    const int slot = slot_search(d2, sT, n_bins);
                    const unsigned fs = in ? ((LOWER ? u : u - fc.cbits) >> fc.k)
                                           : (unsigned)(n_bins + 1);
                    const bool counted = slot >= 1 && slot <= n_bins;
                    if (in && unc) ++audit_unc;
                    if (!(in && unc) && (unsigned)slot != fs &&
                        (counted || (fs >= 1u && fs <= (unsigned)n_bins)))
                        ++audit_bad;
                }
            }
            if (min(u0 & fmask, u1 & fmask) < ff.wlim) {
                const unsigned vb = (v0 ? 1u : 0u) | (v1 ? 2u : 0u);
                if (vb) {
                    const unsigned entry = ((unsigned)tid << 24) | (vb << 22) | (unsigned)q;
                    const unsigned idx = atomicAdd(sCount, 1u);
                    if (idx < (unsigned)kCellListCap) sList[idx] = entry;
                    else cells_filter_fix<EXCL, LOWER>(P, frame, entry, sT, hist32, weight);
                }
            }
        }
    };
    auto wrap = [](int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); };
    auto sweep_row = [&](int y, int z) {
        const int row = (z * g.nc[1] + y) * g.nc[0];
        const int xa = max(cx - 1, 0), xb = min(cx + 1, g.nc[0] - 1);
        const int b0 = start[row + xa];
        sweep(b0, valid ? start[row + xb + 1] - b0 : 0);
        const int xw = (cx == 0) ? g.nc[0] - 1 : (cx == g.nc[0] - 1 ? 0 : -1);
        const int bw = xw >= 0 ? start[row + xw] : 0;
        sweep(bw, (valid && xw >= 0) ? start[row + xw + 1] - bw : 0);
    };

    if (!P.half) {
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                sweep_row(wrap(cy + dy, g.nc[1]), wrap(cz + dz, g.nc[2]));
    } else {
        // half stencil (see rdf_cells_kernel); the own cell contributes j > i with
        // weight 2 and the self pair (distance 0) once
        for (int dy = -1; dy <= 1; ++dy)
            sweep_row(wrap(cy + dy, g.nc[1]), wrap(cz + 1, g.nc[2]));
        sweep_row(wrap(cy + 1, g.nc[1]), cz);
        const int row = (cz * g.nc[1] + cy) * g.nc[0];
        const int xr = wrap(cx + 1, g.nc[0]);
        const int br = start[row + xr];
        sweep(br, valid ? start[row + xr + 1] - br : 0);
        sweep(i + 1, valid ? start[row + cx + 1] - (i + 1) : 0);
        if (valid) {
            my_evals += 1;
            const int slot = slot_search(0.0, sT, n_bins);
            if (!(EXCL) && slot >= 1 && slot <= n_bins)
                red_shared(hist32 + 4u * ((unsigned)slot << fc.sb), 1u);
        }
    }
    __syncthreads();

    // exact re-evaluation of the uncertain pairs of this block
    const unsigned n_push = *sCount;
    const unsigned n_list = min(n_push, (unsigned)kCellListCap);
    for (unsigned e = tid; e < n_list; e += kCfThreads)
        cells_filter_fix<EXCL, LOWER>(P, frame, sList[e], sT, hist32, weight);
    if (tid == 0 && n_push) {
        atomicAdd(&P.fstats[0], (unsigned long long)n_list);
        if (n_push > n_list) atomicAdd(&P.fstats[1], (unsigned long long)(n_push - n_list));
    }
    __syncthreads();

    for (int k = tid; k < n_bins; k += kCfThreads) {
        unsigned sum = 0;                 // modulo 2^32 across warps, see rdf_filter.cu
        for (int w = 0; w < kCfWarps; ++w)
            for (int q = 0; q < (1 << fc.sb); ++q)
                sum += sH[w * hwords + ((k + 1) << fc.sb) + q];
        if (sum) atomicAdd(&P.counts[k], (unsigned long long)sum);
    }
    for (int o = 16; o; o >>= 1) my_evals += __shfl_xor_sync(0xffffffffu, my_evals, o);
    if (lane == 0 && my_evals) atomicAdd(P.evals, my_evals);
    if (AUDIT) {
        if (audit_bad) atomicAdd(&P.fstats[2], audit_bad);
        if (audit_unc) atomicAdd(&P.fstats[3], audit_unc);
    }
}

template <bool EXCL, bool LOWER, bool AUDIT>
int launch_cells_filter_t(mdh_ctx *c, const CellParams &P, dim3 grid)
{
    const size_t smem = cells_filter_smem_bytes(P.n_bins, P.fc.sb);
    auto kern = rdf_cells_filter_kernel<EXCL, LOWER, AUDIT>;
    MDH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    kern<<<dim3((unsigned)((P.n1 + kCfThreads - 1) / kCfThreads), grid.y), kCfThreads, smem,
         c->stream>>>(P);
    MDH_CUDA(cudaGetLastError());
    c->launches++;
    return MDH_OK;
}

template <bool EXCL>
int launch_cells_filter(mdh_ctx *c, const CellParams &P, dim3 grid, bool audit)
{
    if (P.fc.lower)
        return audit ? launch_cells_filter_t<EXCL, true, true>(c, P, grid)
                     : launch_cells_filter_t<EXCL, true, false>(c, P, grid);
    return audit ? launch_cells_filter_t<EXCL, false, true>(c, P, grid)
                 : launch_cells_filter_t<EXCL, false, false>(c, P, grid);
}

template <int HIST, bool EXCL, bool FAST>
int launch_cells(mdh_ctx *c, const CellParams &P, dim3 grid)
{
    const size_t smem = cells_smem_bytes<HIST>(P.n_bins);
    auto kern = rdf_cells_kernel<HIST, EXCL, FAST>;
    MDH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    kern<<<grid, kThreads, smem, c->stream>>>(P);
    MDH_CUDA(cudaGetLastError());
    c->launches++;
    return MDH_OK;
}

int sort_group(mdh_ctx *c, const float4 *pk, int64_t npad, int n, int n_frames,
               const CellGrid *grids, int cstride, DevBuf &cnt, DevBuf &start, DevBuf &rank,
               DevBuf &sorted, DevBuf *pairs)
{
    const int npair = (n + 1) / 2;
    if (pairs)
        if (int rc = pairs->reserve(sizeof(float) * 8 * (size_t)npair * n_frames)) return rc;
    if (int rc = cnt.reserve(sizeof(int) * (size_t)cstride * n_frames)) return rc;
    if (int rc = start.reserve(sizeof(int) * (size_t)cstride * n_frames)) return rc;
    if (int rc = rank.reserve(sizeof(int) * (size_t)n * n_frames)) return rc;
    if (int rc = sorted.reserve(sizeof(float4) * (size_t)n * n_frames)) return rc;
    MDH_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int) * (size_t)cstride * n_frames, c->stream));
    dim3 grid((unsigned)std::min((n + 255) / 256, 2048), n_frames);
    cells_count_kernel<<<grid, 256, 0, c->stream>>>(pk, npad, n, grids, cnt.as<int>(), cstride,
                                                    rank.as<int>());
    MDH_CUDA(cudaGetLastError());
    cells_scan_kernel<<<n_frames, 1024, 0, c->stream>>>(cnt.as<int>(), start.as<int>(), cstride,
                                                        grids);
    MDH_CUDA(cudaGetLastError());
    cells_scatter_kernel<<<grid, 256, 0, c->stream>>>(pk, npad, n, grids, start.as<int>(),
                                                      cstride, rank.as<int>(),
                                                      sorted.as<float4>(),
                                                      pairs ? pairs->as<float>() : nullptr, npair);
    MDH_CUDA(cudaGetLastError());
    c->launches += 3;
    return MDH_OK;
}

}  // namespace

// Called from rdf_accumulate_impl after the packed float4 arrays and the FrameBox
// array of the batch are on the device.
int rdf_cells_accumulate(mdh_ctx *c, int f0, int n_frames, bool use_filter)
{
    RdfState &R = c->rdf;
    // the filter kernel addresses neighbour pairs with 22 bits
    use_filter = use_filter && (R.n2 + 1) / 2 <= (1 << 22) &&
                 cells_filter_smem_bytes(R.n_bins, R.fc.sb) <= 100 * 1024;
    MDH_REQUIRE(R.drop_axis < 0, MDH_EINVAL, "rdf: cell-list mode does not support drop_axis");
    const double r_cut = sqrt(R.thr_hi) * 1.00001;
    std::vector<CellGrid> grids(n_frames);
    int ncell_max = 0;
    for (int f = 0; f < n_frames; ++f) {
        CellGrid &g = grids[f];
        double ncell = 1;
        for (int k = 0; k < 3; ++k) {
            g.box[k] = R.h_boxes[f0 + f].box[k];
            int nc = (int)floor(g.box[k] / r_cut);
            MDH_REQUIRE(nc >= 3, MDH_EINVAL,
                        "rdf: cell-list mode needs box edge >= 3*r_max (frame %d axis %d)", f, k);
            g.nc[k] = std::min(nc, 160);
            ncell *= g.nc[k];
        }
        // keep at least ~2 particles per cell on average
        const double want = std::max(27.0, (double)std::max(R.n1, R.n2) / 2.0);
        while (ncell > want) {
            int kmax = 0;
            for (int k = 1; k < 3; ++k) if (g.nc[k] > g.nc[kmax]) kmax = k;
            if (g.nc[kmax] <= 3) break;
            ncell = ncell / g.nc[kmax] * (g.nc[kmax] - 1);
            g.nc[kmax]--;
        }
        for (int k = 0; k < 3; ++k) g.inv_w[k] = g.nc[k] / g.box[k];
        g.ncell = g.nc[0] * g.nc[1] * g.nc[2];
        ncell_max = std::max(ncell_max, g.ncell);
    }
    const int cstride = ncell_max + 1;
    DevBuf &d_grids = R.cell[0];
    if (int rc = d_grids.reserve(sizeof(CellGrid) * n_frames)) return rc;
    // pageable source: the runtime stages it before returning, so the local vector
    // may go out of scope while the copy is still queued
    MDH_CUDA(cudaMemcpyAsync(d_grids.p, grids.data(), sizeof(CellGrid) * n_frames,
                             cudaMemcpyHostToDevice, c->stream));

    const int tile = kThreads * R.ipt;
    const int64_t pad1 = (R.n1 + tile - 1) / tile * tile;
    const int64_t pad2 = (R.n2 + tile - 1) / tile * tile;
    if (int rc = sort_group(c, R.pk1.as<float4>(), pad1, (int)R.n1, n_frames,
                            d_grids.as<CellGrid>(), cstride, R.cell[1], R.cell[2], R.cell[3],
                            R.cell[4], (R.same && use_filter) ? &R.cell_pairs : nullptr))
        return rc;
    if (!R.same)
        if (int rc = sort_group(c, R.pk2.as<float4>(), pad2, (int)R.n2, n_frames,
                                d_grids.as<CellGrid>(), cstride, R.cell[5], R.cell[6],
                                R.cell[7], R.cell[8], use_filter ? &R.cell_pairs : nullptr))
            return rc;

    CellParams P;
    P.s1 = R.cell[4].as<float4>();
    P.s2 = R.same ? P.s1 : R.cell[8].as<float4>();
    P.n1 = (int)R.n1; P.n2 = (int)R.n2;
    P.start2 = R.same ? R.cell[2].as<int>() : R.cell[6].as<int>();
    P.cstride = cstride;
    P.grids = d_grids.as<CellGrid>();
    P.boxes = R.boxes.as<FrameBox>() + f0;
    P.thr = R.thr.as<double>();
    P.n_bins = R.n_bins;
    P.guess = rdf_bin_guess(R);
    P.counts = R.counts.as<unsigned long long>();
    P.evals = R.cell[9].as<unsigned long long>();
    P.half = R.same;
    P.pairs2 = nullptr; P.npair2 = (int)((R.n2 + 1) / 2);
    P.filt = nullptr;
    P.fc = R.fc;
    P.fast_bins = R.fast_bins ? 1 : 0;
    P.fstats = R.fstats.as<unsigned long long>();
    dim3 grid((unsigned)((R.n1 + kThreads - 1) / kThreads), (unsigned)n_frames);
    const bool excl = R.excl1 > 0, fast = R.fast_bins;
    if (use_filter) {
        // the filter kernel takes the frames whose error bound is small against a bin;
        // the exact kernel below only runs the ones it declined
        P.pairs2 = R.cell_pairs.as<float4>();
        P.filt = R.filt.as<FrameFilter>();
        const bool audit = R.filter_mode == MDH_FILTER_AUDIT;
        if (int rc = excl ? launch_cells_filter<true>(c, P, grid, audit)
                          : launch_cells_filter<false>(c, P, grid, audit)) return rc;
    }
    if (R.hist == MDH_HIST_LANE_PRIVATE) {
        if (excl)
            return fast ? launch_cells<MDH_HIST_LANE_PRIVATE, true, true>(c, P, grid)
                        : launch_cells<MDH_HIST_LANE_PRIVATE, true, false>(c, P, grid);
        return fast ? launch_cells<MDH_HIST_LANE_PRIVATE, false, true>(c, P, grid)
                    : launch_cells<MDH_HIST_LANE_PRIVATE, false, false>(c, P, grid);
    }
    if (excl)
        return fast ? launch_cells<MDH_HIST_WARP_ATOMIC, true, true>(c, P, grid)
                    : launch_cells<MDH_HIST_WARP_ATOMIC, true, false>(c, P, grid);
    return fast ? launch_cells<MDH_HIST_WARP_ATOMIC, false, true>(c, P, grid)
                : launch_cells<MDH_HIST_WARP_ATOMIC, false, false>(c, P, grid);
}
