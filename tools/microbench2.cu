// microbench2.cu -- issue-side rates that bound the fp32-filter pair kernel: packed
// f32x2 arithmetic (FFMA2 / FADD2), the ALU-pipe integer ops of the bin stage, and
// shared-memory RED throughput without loop overhead.  Prints one JSON object.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench2 tools/microbench2.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ITERS = 16384;
constexpr int ILP = 8;

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b)
{
    uint64_t d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t pack(float lo, float hi)
{
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}

#define KERNEL(name, T, INIT, OP)                                                   \
    __global__ void name(unsigned long long *out, float a, float b) {               \
        T v[ILP];                                                                   \
        for (int i = 0; i < ILP; ++i) v[i] = INIT;                                  \
        for (int it = 0; it < ITERS; ++it) {                                        \
            _Pragma("unroll") for (int i = 0; i < ILP; ++i) { OP; }                 \
        }                                                                           \
        unsigned long long s = 0;                                                   \
        for (int i = 0; i < ILP; ++i) s += (unsigned long long)v[i];                \
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;                             \
    }

KERNEL(k_ffma, float, a + i + threadIdx.x, v[i] = fmaf(v[i], a, b))
KERNEL(k_fadd, float, a + i + threadIdx.x, v[i] = __fadd_rn(v[i], b))
KERNEL(k_ffma2, uint64_t, pack(a + i, b + threadIdx.x), v[i] = ffma2(v[i], pack(a, a), pack(b, b)))
KERNEL(k_fadd2, uint64_t, pack(a + i, b + threadIdx.x), v[i] = fadd2(v[i], pack(b, b)))
KERNEL(k_imad, unsigned, i + threadIdx.x, v[i] = v[i] * (unsigned)a + (unsigned)b)
KERNEL(k_vimnmx, unsigned, i + threadIdx.x * 77, v[i] = min(v[i] ^ 0x55u, (unsigned)b + it))
KERNEL(k_shf, unsigned, i + threadIdx.x * 77, v[i] = (v[i] >> ((unsigned)b & 7)) + 0x10000000u)
// FFMA and ALU ops interleaved 2:1 -- do the two pipes issue side by side?
KERNEL(k_mix, float, a + i + threadIdx.x,
       v[i] = fmaf(v[i], a, b); v[i] = fmaf(v[i], a, b);
       v[i] = __uint_as_float(min(__float_as_uint(v[i]), 0x4f000000u + it)))

// DFMA with three distinct vector-register operands, the S(q) inner loop's shape:
// 32 accumulators, acc[m][r] += a[m] * z[r] (re/im mixes), operands refreshed from
// registers only.  Counts DFMAs.
__global__ void k_dfma3(unsigned long long *out, double a0, double b0)
{
    double acc_re[2][8], acc_im[2][8], ar[2], ai[2], zr[8], zi[8];
    for (int m = 0; m < 2; ++m) { ar[m] = a0 + m + threadIdx.x; ai[m] = b0 + m; }
    for (int r = 0; r < 8; ++r) { zr[r] = a0 * r + 1; zi[r] = b0 * r + threadIdx.x; }
    for (int m = 0; m < 2; ++m)
        for (int r = 0; r < 8; ++r) acc_re[m][r] = acc_im[m][r] = 0;
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                acc_re[m][r] = fma(ar[m], zr[r], acc_re[m][r]);
                acc_re[m][r] = fma(-ai[m], zi[r], acc_re[m][r]);
                acc_im[m][r] = fma(ar[m], zi[r], acc_im[m][r]);
                acc_im[m][r] = fma(ai[m], zr[r], acc_im[m][r]);
            }
        // keep the compiler from hoisting: rotate the operands through the accumulators
        ar[0] += acc_im[1][7] * 1e-300; zr[0] += acc_re[0][0] * 1e-300;
    }
    double s = 0;
    for (int m = 0; m < 2; ++m)
        for (int r = 0; r < 8; ++r) s += acc_re[m][r] + acc_im[m][r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (unsigned long long)s;
}

// RED.shared with per-thread precomputed word indices (ILP addresses in registers)
template <int MODE>   // 0: every lane its own bank, 1: random words of a 201*4-word histogram
__global__ void k_red(unsigned long long *out, int n_words)
{
    extern __shared__ unsigned sh[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (blockDim.x / 32) * n_words; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    unsigned addr[ILP];
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    for (int i = 0; i < ILP; ++i) {
        x = x * 1664525u + 1013904223u;
        const unsigned w = MODE == 0 ? (unsigned)(lane + 32 * i) % n_words : (x >> 8) % n_words;
        addr[i] = (unsigned)__cvta_generic_to_shared(sh + warp * n_words + w);
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr[i]), "r"(1u));
    }
    __syncthreads();
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < (blockDim.x / 32) * n_words; i += blockDim.x) s += sh[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char *name, F launch, double ops, int sms, int khz, bool last = false)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 6; ++w) launch();
    CK(cudaDeviceSynchronize());
    const int reps = 4;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double gops = ops / (ms * 1e-3) / 1e9;
    printf("  \"%s\": {\"gops_per_s\": %.1f, \"per_clk_per_sm_at_max_clock\": %.2f, \"ms\": %.4f}%s\n",
           name, gops, gops * 1e9 / ((double)sms * khz * 1e3), ms, last ? "" : ",");
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount, khz = p.clockRate;
    const int bps = 2, threads = 512, blocks = sms * bps;
    unsigned long long *out;
    CK(cudaMalloc(&out, sizeof(*out) * blocks * threads));
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"note\": \"thread-level "
           "instructions per second (one f32x2 instruction counts once)\",\n", p.name, sms, khz);
    const double n = (double)ITERS * ILP * threads * blocks;
    run("ffma", [&] { k_ffma<<<blocks, threads>>>(out, 1.0000001f, 1e-9f); }, n, sms, khz);
    run("fadd", [&] { k_fadd<<<blocks, threads>>>(out, 1.0000001f, 1e-9f); }, n, sms, khz);
    run("ffma2", [&] { k_ffma2<<<blocks, threads>>>(out, 1.0000001f, 1e-9f); }, n, sms, khz);
    run("fadd2", [&] { k_fadd2<<<blocks, threads>>>(out, 1.0000001f, 1e-9f); }, n, sms, khz);
    run("dfma_3reg_operands", [&] { k_dfma3<<<blocks, 256>>>(out, 1.0000001, 1e-9); },
        (double)(ITERS / 8) * 64 * 256 * blocks, sms, khz);
    run("imad", [&] { k_imad<<<blocks, threads>>>(out, 3.f, 7.f); }, n, sms, khz);
    run("vimnmx", [&] { k_vimnmx<<<blocks, threads>>>(out, 3.f, 7.f); }, n, sms, khz);
    run("shf_iadd", [&] { k_shf<<<blocks, threads>>>(out, 3.f, 7.f); }, 2 * n, sms, khz);
    run("mix_2ffma_1vimnmx", [&] { k_mix<<<blocks, threads>>>(out, 1.0000001f, 1e-9f); }, 3 * n, sms, khz);
    const int nw = 201 * 4 + 32;
    const size_t smem = sizeof(unsigned) * (threads / 32) * nw;
    CK(cudaFuncSetAttribute(k_red<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_red<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    run("red_shared_conflict_free", [&] { k_red<0><<<blocks, threads, smem>>>(out, nw); }, n, sms, khz);
    run("red_shared_random_words", [&] { k_red<1><<<blocks, threads, smem>>>(out, nw); }, n, sms, khz, true);
    printf("}\n");
    return 0;
}
