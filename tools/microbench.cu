// microbench.cu -- measured per-SM pipe rates on the box the bench runs on
// (SURVEY.md section 7 step 0): the denominators of the pair-kernel (FP64 pipe)
// and S(q) (FP64 FMA / SFU) rooflines, plus shared-memory histogram update rates.
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ITERS = 32768;
constexpr int ILP = 8;

struct Cyc { unsigned long long c; };

#define PIPE_KERNEL(name, T, INIT, OP)                                              \
    __global__ void name(T *out, Cyc *cyc, T a, T b) {                              \
        T v[ILP];                                                                   \
        for (int i = 0; i < ILP; ++i) v[i] = INIT;                                  \
        __syncthreads();                                                            \
        unsigned long long t0 = clock64();                                          \
        for (int it = 0; it < ITERS; ++it) {                                        \
            _Pragma("unroll") for (int i = 0; i < ILP; ++i) { OP; }                 \
        }                                                                           \
        unsigned long long t1 = clock64();                                          \
        T s = v[0];                                                                 \
        for (int i = 1; i < ILP; ++i) s += v[i];                                    \
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;                             \
        if (threadIdx.x == 0) cyc[blockIdx.x].c = t1 - t0;                          \
    }

PIPE_KERNEL(k_dfma, double, a + i + threadIdx.x, v[i] = fma(v[i], a, b))
PIPE_KERNEL(k_dmul, double, a + i + threadIdx.x, v[i] = __dmul_rn(v[i], a))
PIPE_KERNEL(k_dadd, double, a + i + threadIdx.x, v[i] = __dadd_rn(v[i], b))
PIPE_KERNEL(k_ffma, float, a + i + threadIdx.x, v[i] = fmaf(v[i], a, b))
PIPE_KERNEL(k_sin, float, a + i + threadIdx.x, v[i] = __sinf(v[i]))
PIPE_KERNEL(k_sqrt, float, a + i + threadIdx.x,
            asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(v[i])))

__global__ void k_f2d(double *out, Cyc *cyc, float a, float b)
{
    float v[ILP];
    double acc[ILP];
    for (int i = 0; i < ILP; ++i) { v[i] = a + i + threadIdx.x; acc[i] = 0; }
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            double d = (double)v[i];                       // F2F.F64.F32
            v[i] = __int_as_float(__double2hiint(d) ^ it); // cheap ALU feedback
            acc[i] = d;
        }
    }
    unsigned long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += acc[i] + v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x].c = t1 - t0;
}

// shared-memory histogram update variants; 201 bins, bins from a cheap LCG
enum { H_WARP_ATOMIC = 0, H_LANE_ATOMIC = 1, H_LANE_RMW = 2, H_BLOCK_ATOMIC = 3 };
template <int MODE>
__global__ void k_hist(unsigned *out, Cyc *cyc, int n_bins)
{
    extern __shared__ unsigned sh[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_words = (n_bins + 3) / 4;
    const int total = MODE == H_WARP_ATOMIC ? (blockDim.x / 32) * n_bins
                    : MODE == H_BLOCK_ATOMIC ? n_bins
                    : (blockDim.x / 32) * n_words * 32;
    for (int i = threadIdx.x; i < total; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    unsigned long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        x = x * 1664525u + 1013904223u;
        // r^2-weighted bins, like pair distances in a liquid
        float u = (x >> 8) * (1.0f / 16777216.0f);
        int k = (int)(cbrtf(u) * n_bins);
        k = min(k, n_bins - 1);
        if (MODE == H_WARP_ATOMIC) atomicAdd(&sh[warp * n_bins + k], 1u);
        else if (MODE == H_BLOCK_ATOMIC) atomicAdd(&sh[k], 1u);
        else if (MODE == H_LANE_ATOMIC)
            atomicAdd(&sh[(warp * n_words + (k >> 2)) * 32 + lane], 1u << ((k & 3) * 8));
        else {
            unsigned *w = &sh[(warp * n_words + (k >> 2)) * 32 + lane];
            *w += 1u << ((k & 3) * 8);
            if ((it & 127) == 127) *w = 0;     // keep the bytes from overflowing
        }
    }
    unsigned long long t1 = clock64();
    __syncthreads();
    unsigned s = 0;
    for (int i = threadIdx.x; i < total; i += blockDim.x) s += sh[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + x;
    if (threadIdx.x == 0) cyc[blockIdx.x].c = t1 - t0;
}

// same loop without the histogram update: subtract to isolate the update cost
__global__ void k_hist_base(unsigned *out, Cyc *cyc, int n_bins)
{
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    unsigned acc = 0;
    unsigned long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        x = x * 1664525u + 1013904223u;
        float u = (x >> 8) * (1.0f / 16777216.0f);
        int k = (int)(cbrtf(u) * n_bins);
        acc += min(k, n_bins - 1);
    }
    unsigned long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x].c = t1 - t0;
}

template <typename F>
static void run(const char *name, F launch, int blocks, int threads, double ops_per_thread,
                int blocks_per_sm, Cyc *d_cyc, bool last = false)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 8; ++w) launch();       // warm-up: let the clocks ramp
    CK(cudaDeviceSynchronize());
    const int reps = 4;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    Cyc *h = (Cyc *)malloc(sizeof(Cyc) * blocks);
    CK(cudaMemcpy(h, d_cyc, sizeof(Cyc) * blocks, cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < blocks; ++i) mean += (double)h[i].c;
    mean /= blocks;
    free(h);
    // all resident blocks of an SM run concurrently for ~mean cycles
    const double per_clk_sm = ops_per_thread * threads * blocks_per_sm / mean;
    const double mhz = mean / (ms * 1e3);       // clock64 ticks per block / wall time
    const double gops = ops_per_thread * threads * blocks / (ms * 1e-3) / 1e9;   // whole GPU
    printf("  \"%s\": {\"gops_per_s\": %.1f, \"ops_per_clk_per_sm\": %.2f, \"ms\": %.4f, "
           "\"clock64_mhz\": %.0f}%s\n", name, gops, per_clk_sm, ms, mhz, last ? "" : ",");
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const int bps = 2, threads = 512, blocks = sms * bps;
    void *out; Cyc *cyc;
    CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
    CK(cudaMalloc(&cyc, sizeof(Cyc) * blocks));
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, sms, p.clockRate);
    const double n = (double)ITERS * ILP;
    run("dfma", [&] { k_dfma<<<blocks, threads>>>((double *)out, cyc, 1.0000001, 1e-9); }, blocks, threads, n, bps, cyc);
    run("dmul", [&] { k_dmul<<<blocks, threads>>>((double *)out, cyc, 1.0000001, 1e-9); }, blocks, threads, n, bps, cyc);
    run("dadd", [&] { k_dadd<<<blocks, threads>>>((double *)out, cyc, 1.0000001, 1e-9); }, blocks, threads, n, bps, cyc);
    run("ffma", [&] { k_ffma<<<blocks, threads>>>((float *)out, cyc, 1.0000001f, 1e-9f); }, blocks, threads, n, bps, cyc);
    run("mufu_sin", [&] { k_sin<<<blocks, threads>>>((float *)out, cyc, 0.5f, 0.f); }, blocks, threads, n, bps, cyc);
    run("mufu_sqrt", [&] { k_sqrt<<<blocks, threads>>>((float *)out, cyc, 0.5f, 0.f); }, blocks, threads, n, bps, cyc);
    run("f2f_f64_f32", [&] { k_f2d<<<blocks, threads>>>((double *)out, cyc, 0.5f, 0.f); }, blocks, threads, n, bps, cyc);
    const int nb = 201, nw = (nb + 3) / 4, warps = threads / 32;
    const size_t s_warp = sizeof(unsigned) * warps * nb, s_lane = sizeof(unsigned) * warps * nw * 32;
    CK(cudaFuncSetAttribute(k_hist<H_LANE_ATOMIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s_lane));
    CK(cudaFuncSetAttribute(k_hist<H_LANE_RMW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s_lane));
    run("hist_base_loop", [&] { k_hist_base<<<blocks, threads>>>((unsigned *)out, cyc, nb); }, blocks, threads, ITERS, bps, cyc);
    run("hist_block_atomic", [&] { k_hist<H_BLOCK_ATOMIC><<<blocks, threads, sizeof(unsigned) * nb>>>((unsigned *)out, cyc, nb); }, blocks, threads, ITERS, bps, cyc);
    run("hist_warp_atomic", [&] { k_hist<H_WARP_ATOMIC><<<blocks, threads, s_warp>>>((unsigned *)out, cyc, nb); }, blocks, threads, ITERS, bps, cyc);
    int occ_a = 1, occ_r = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_a, k_hist<H_LANE_ATOMIC>, threads, s_lane));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_r, k_hist<H_LANE_RMW>, threads, s_lane));
    run("hist_lane_atomic", [&] { k_hist<H_LANE_ATOMIC><<<blocks, threads, s_lane>>>((unsigned *)out, cyc, nb); }, blocks, threads, ITERS, occ_a < bps ? occ_a : bps, cyc);
    run("hist_lane_rmw", [&] { k_hist<H_LANE_RMW><<<blocks, threads, s_lane>>>((unsigned *)out, cyc, nb); }, blocks, threads, ITERS, occ_r < bps ? occ_r : bps, cyc, true);
    printf("}\n");
    return 0;
}
