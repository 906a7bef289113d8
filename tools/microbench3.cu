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
        warps * ITERS * 8 * 2048, sms, khz, true);
    printf("}\n");
    return 0;
}
