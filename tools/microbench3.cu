// microbench3.cu -- FP64 matrix-multiply-accumulate (mma.sync ... f64, SASS DMMA) rates on
// sm_100a, in scalar-FMA equivalents per clock and SM, next to the plain DFMA figure of
// microbench.cu / microbench2.cu.  The S(q) lattice sum is a complex rank-K update
// rho[(nx,ny)][nz] += A[(nx,ny)][j] * E_z[j][nz]; this measures what the tensor path
// could sustain for it.  Prints one JSON object.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench3 tools/microbench3.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ITERS = 4096;

// m8n8k4: A 8x4 (1 reg/thread), B 4x8 (1 reg/thread), C 8x8 (2 regs/thread); 256 FMA
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// m16n8k4: A 16x4 (2), B 4x8 (1), C 16x8 (4); 512 FMA
__device__ __forceinline__ void dmma1684(double *c, const double *a, double b)
{
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
// m16n8k8: A 16x8 (4), B 8x8 (2), C 16x8 (4); 1024 FMA
__device__ __forceinline__ void dmma1688(double *c, const double *a, const double *b)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
// m16n8k16: A 16x16 (8), B 16x8 (4), C 16x8 (4); 2048 FMA
__device__ __forceinline__ void dmma16816(double *c, const double *a, const double *b)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                 "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_884(double *out, double a0, double b0)
{
    double c[NACC][2];
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0;
    double a = a0 + threadIdx.x, b = b0 * threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_1684(double *out, double a0, double b0)
{
    double c[NACC][4], a[2] = {a0 + threadIdx.x, a0 - threadIdx.x}, b = b0 * threadIdx.x;
    for (int i = 0; i < NACC; ++i) for (int k = 0; k < 4; ++k) c[i][k] = 0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1684(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) for (int k = 0; k < 4; ++k) s += c[i][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_1688(double *out, double a0, double b0)
{
    double c[NACC][4], a[4], b[2];
    for (int k = 0; k < 4; ++k) a[k] = a0 + k * threadIdx.x;
    for (int k = 0; k < 2; ++k) b[k] = b0 * (k + threadIdx.x);
    for (int i = 0; i < NACC; ++i) for (int k = 0; k < 4; ++k) c[i][k] = 0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1688(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) for (int k = 0; k < 4; ++k) s += c[i][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_16816(double *out, double a0, double b0)
{
    double c[NACC][4], a[8], b[4];
    for (int k = 0; k < 8; ++k) a[k] = a0 + k * threadIdx.x;
    for (int k = 0; k < 4; ++k) b[k] = b0 * (k + threadIdx.x);
    for (int i = 0; i < NACC; ++i) for (int k = 0; k < 4; ++k) c[i][k] = 0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma16816(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) for (int k = 0; k < 4; ++k) s += c[i][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// DMMA bursts separated by scalar FP64 work, the shape of the S(q) consumer loop: ND DMMAs,
// then NF dependent DFMAs whose result is (FEED) or is not the next burst's A operand
template <int ND, int NF, bool FEED>
__global__ void k_mix(double *out, double a0, double b0)
{
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0;
    double a = a0 + threadIdx.x, b = b0 * threadIdx.x, f = a0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ND; ++i) dmma884(c[i & 7][0], c[i & 7][1], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) f = fma(f, a0, b0);
        if (FEED) a = f;
    }
    double s = f;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// same with the scalar work on the FP32 pipe
template <int ND, int NF>
__global__ void k_mix32(double *out, double a0, double b0)
{
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0;
    double a = a0 + threadIdx.x, b = b0 * threadIdx.x;
    float f = (float)a0, fa = (float)a0, fb = (float)b0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ND; ++i) dmma884(c[i & 7][0], c[i & 7][1], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) f = fmaf(f, fa, fb);
        a = __hiloint2double(__float_as_int(f), __double2loint(a));
    }
    double s = f;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the consumer loop's register pattern: 16 accumulators (8 re, 8 im), two A values, four B
// pairs, two passes per iteration
__global__ void k_pattern(double *out, double a0, double b0)
{
    double cre[8][2], cim[8][2];
    for (int i = 0; i < 8; ++i) cre[i][0] = cre[i][1] = cim[i][0] = cim[i][1] = 0;
    double ar[2], ai[2], nai[2], zx[4], zy[4];
    for (int i = 0; i < 2; ++i) { ar[i] = a0 + i + threadIdx.x; ai[i] = a0 * (i + 2); nai[i] = -ai[i]; }
    for (int t = 0; t < 4; ++t) { zx[t] = b0 * (t + 1) * threadIdx.x; zy[t] = b0 + t; }
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                dmma884(cre[i * 4 + t][0], cre[i * 4 + t][1], ar[i], zx[t]);
                dmma884(cim[i * 4 + t][0], cim[i * 4 + t][1], ar[i], zy[t]);
            }
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                dmma884(cre[i * 4 + t][0], cre[i * 4 + t][1], nai[i], zy[t]);
                dmma884(cim[i * 4 + t][0], cim[i * 4 + t][1], ai[i], zx[t]);
            }
        ar[0] += 1e-300; zx[0] += 1e-300;
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += cre[i][0] + cre[i][1] + cim[i][0] + cim[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// does other work issue beside a DMMA?  8 DMMAs per iteration, NI integer IMADs (or NL
// shared-memory loads) per DMMA in between
template <int NI, int NL>
__global__ void k_fill(double *out, double a0, double b0)
{
    __shared__ double2 sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = make_double2(a0 + i, b0 * i);
    __syncthreads();
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0;
    double a = a0 + threadIdx.x, b = b0 * threadIdx.x;
    unsigned x = threadIdx.x, y = blockIdx.x + 1;
    double2 acc = make_double2(0, 0);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
            for (int j = 0; j < NI; ++j) x = x * y + 12345u;
#pragma unroll
            for (int j = 0; j < NL; ++j) {
                const double2 v = sh[(threadIdx.x + 37 * (i * NL + j) + it) & 1023];
                acc.x += v.x;
            }
        }
    }
    double s = x + acc.x;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// which warps share a scheduler?  only the warps of `mask` (bit = warp index in the block)
// issue DMMAs; also records %warpid of every warp
__global__ void k_mask(double *out, double a0, double b0, unsigned mask, unsigned *slots)
{
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0;
    double a = a0 + threadIdx.x, b = b0 * threadIdx.x;
    const int w = threadIdx.x >> 5;
    unsigned hw;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw));
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) slots[w] = hw;
    if ((mask >> w) & 1) {
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// plain DFMA, two uniform operands (the figure used as nominal peak so far)
__global__ void k_dfma(double *out, double a, double b)
{
    double v[8];
    for (int i = 0; i < 8; ++i) v[i] = a + i + threadIdx.x;
    for (int it = 0; it < ITERS * 4; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fma(v[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char *name, F launch, double fmas, int sms, int khz, bool last = false)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 4; ++w) launch();
    CK(cudaDeviceSynchronize());
    const int reps = 4;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double g = fmas / (ms * 1e-3) / 1e9;
    printf("  \"%s\": {\"gfma_per_s\": %.1f, \"fma_per_clk_per_sm_at_max_clock\": %.2f, \"ms\": %.4f}%s\n",
           name, g, g * 1e9 / ((double)sms * khz * 1e3), ms, last ? "" : ",");
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount, khz = p.clockRate;
    const int threads = 256, blocks = sms * 4;
    double *out;
    CK(cudaMalloc(&out, sizeof(double) * blocks * 1024));
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"note\": \"scalar fp64 FMA "
           "equivalents (an m8n8k4 DMMA = 256)\",\n", p.name, sms, khz);
    const double warps = (double)blocks * threads / 32;
    run("dfma_uniform_operands", [&] { k_dfma<<<blocks, threads>>>(out, 1.0000001, 1e-9); },
        (double)ITERS * 4 * 8 * threads * blocks, sms, khz);
    run("dmma_m8n8k4_acc1_dependent_chain", [&] { k_884<1><<<blocks, threads>>>(out, 1.0000001, 1e-9); },
        warps * ITERS * 1 * 256, sms, khz);
    run("dmma_m8n8k4_acc2", [&] { k_884<2><<<blocks, threads>>>(out, 1.0000001, 1e-9); },
        warps * ITERS * 2 * 256, sms, khz);
    run("dmma_m8n8k4_acc4", [&] { k_884<4><<<blocks, threads>>>(out, 1.0000001, 1e-9); },
        warps * ITERS * 4 * 256, sms, khz);
    run("dmma_m8n8k4_acc8", [&] { k_884<8><<<blocks, threads>>>(out, 1.0000001, 1e-9); },
        warps * ITERS * 8 * 256, sms, khz);
    run("dmma_m8n8k4_acc16", [&] { k_884<16><<<blocks, threads>>>(out, 1.0000001, 1e-9); },
        warps * ITERS * 16 * 256, sms, khz);
    run("dmma_m16n8k4_acc8", [&] { k_1684<8><<<blocks, threads>>>(out, 1.0000001, 1e-9); },
        warps * ITERS * 8 * 512, sms, khz);
    run("dmma_m16n8k8_acc8", [&] { k_1688<8><<<blocks, threads>>>(out, 1.0000001, 1e-9); },
        warps * ITERS * 8 * 1024, sms, khz);
    run("dmma_m16n8k16_acc4", [&] { k_16816<4><<<blocks, threads>>>(out, 1.0000001, 1e-9); },
        warps * ITERS * 4 * 2048, sms, khz);
    run("dmma_m16n8k16_acc8", [&] { k_16816<8><<<blocks, threads>>>(out, 1.0000001, 1e-9); },
        warps * ITERS * 8 * 2048, sms, khz);
    // how many warps per scheduler does the DMMA pipe need?  one block per SM, w warps per
    // scheduler, 8 (or 2) independent accumulator pairs per warp
    for (int w = 1; w <= 8; ++w) {
        char name[64];
        snprintf(name, sizeof name, "dmma_m8n8k4_acc8_%dwarps_per_scheduler", w);
        run(name, [&] { k_884<8><<<sms, 128 * w>>>(out, 1.0000001, 1e-9); },
            (double)sms * 4 * w * ITERS * 8 * 256, sms, khz);
    }
#define FILL(NI, NL) \
    run("dmma_plus_" #NI "imad_" #NL "lds_per_dmma", \
        [&] { k_fill<NI, NL><<<sms, 512>>>(out, 1.0000001, 1e-9); }, \
        (double)sms * 16 * ITERS * 8 * 256, sms, khz);
    FILL(0, 0) FILL(2, 0) FILL(4, 0) FILL(8, 0) FILL(12, 0) FILL(16, 0) FILL(0, 1)
#undef FILL
    for (int w = 1; w <= 4; ++w) {
        char name[64];
        snprintf(name, sizeof name, "dmma_consumer_register_pattern_%dwarps_per_scheduler", w);
        run(name, [&] { k_pattern<<<sms, 128 * w>>>(out, 1.0000001, 1e-9); },
            (double)sms * 4 * w * (ITERS / 4) * 32 * 256, sms, khz);
    }
    // four active warps out of 16: on four different schedulers or all on one?
    unsigned *slots;
    CK(cudaMalloc(&slots, 64 * sizeof(unsigned)));
    const unsigned masks[] = {0x000f, 0x1111, 0x0033, 0x0505, 0x8421, 0x00ff, 0x3333, 0xffff};
    for (unsigned m : masks) {
        char name[64];
        snprintf(name, sizeof name, "dmma_active_warp_mask_%04x", m);
        run(name, [&] { k_mask<<<sms, 512>>>(out, 1.0000001, 1e-9, m, slots); },
            (double)sms * __builtin_popcount(m) * ITERS * 8 * 256, sms, khz);
    }
    unsigned hslots[16];
    CK(cudaMemcpy(hslots, slots, sizeof hslots, cudaMemcpyDeviceToHost));
    printf("  \"warpid_of_cta_warps\": [");
    for (int i = 0; i < 16; ++i) printf("%u%s", hslots[i], i < 15 ? ", " : "],\n");
    // DMMA bursts + scalar FP64 work in between, 4 warps per scheduler (rates count DMMA only)
#define MIX(ND, NF, FEED) \
    run("dmma_burst" #ND "_dfma" #NF "_feed" #FEED, \
        [&] { k_mix<ND, NF, FEED><<<sms, 512>>>(out, 1.0000001, 1e-9); }, \
        (double)sms * 16 * ITERS * ND * 256, sms, khz);
    MIX(8, 0, false) MIX(8, 2, true) MIX(8, 2, false) MIX(8, 4, true) MIX(16, 2, true)
    MIX(16, 4, true) MIX(32, 2, true) MIX(32, 4, true) MIX(32, 8, true) MIX(64, 8, true)
#undef MIX
    run("dmma_burst8_ffma4_feed", [&] { k_mix32<8, 4><<<sms, 512>>>(out, 1.0000001, 1e-9); },
        (double)sms * 16 * ITERS * 8 * 256, sms, khz);
    run("dmma_burst16_ffma4_feed", [&] { k_mix32<16, 4><<<sms, 512>>>(out, 1.0000001, 1e-9); },
        (double)sms * 16 * ITERS * 16 * 256, sms, khz);
    for (int w = 1; w <= 8; w *= 2) {
        char name[64];
        snprintf(name, sizeof name, "dmma_m8n8k4_acc2_%dwarps_per_scheduler", w);
        run(name, [&] { k_884<2><<<sms, 128 * w>>>(out, 1.0000001, 1e-9); },
            (double)sms * 4 * w * ITERS * 2 * 256, sms, khz, w == 8);
    }
    printf("}\n");
    return 0;
}
