// ubench_dispatch.cu -- what bounds a mix of packed f32x2 and ALU instructions on an sm_100a scheduler?
// 3 CTAs x 256 threads per SM (6 warps per scheduler), every chain independent (16 accumulators per kind), so nothing
// is latency-bound.  Prints scheduler cycles per warp-iteration next to what three models predict:
//   pipes   : max(FMA-pipe cycles, ALU-pipe cycles)      (packed = 2 FMA cycles, scalar FP32 = 1, ALU op = A cycles)
//   dispatch: one dispatch slot per instruction, a packed instruction holds the dispatch port for 2 cycles
#include <cstdio>
#include <cuda_runtime.h>

template <int NP, int NS, int NA, int KIND>   // NP packed FFMA2, NS scalar FFMA, NA ALU ops (KIND 0: PRMT, 1: LOP3, 2: FSEL-like FMNMX) per iteration
__global__ void __launch_bounds__(256, 3) k(float *out, int iters, float a, float one, unsigned m)
{
    float2 acc2[16]; float acc[16]; unsigned y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc2[i] = make_float2(threadIdx.x * 1e-3f + i, i); acc[i] = threadIdx.x * 1e-3f - i; y[i] = threadIdx.x * 2654435761u + i * 40503u; }
    const float2 a2 = make_float2(a, a), one2 = make_float2(one, one);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (i < NP) acc2[i & 15] = __ffma2_rn(acc2[i & 15], a2, one2);
            if (i < NS) acc[i & 15] = __fmaf_rn(acc[i & 15], a, one);
            if (i < NA) {
                if (KIND == 0) y[i & 15] = __byte_perm(y[i & 15], m, 0x7440 | (i & 3));
                else if (KIND == 1) y[i & 15] = (y[i & 15] ^ m) & (y[i & 15] | 0x55u + i);
                else y[i & 15] = __float_as_uint(fminf(__uint_as_float(y[i & 15]), __uint_as_float(m + i)));
            }
        }
    }
    float s = 0; unsigned u = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { s += acc2[i].x + acc2[i].y + acc[i]; u ^= y[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float) u;
}

template <int NP, int NS, int NA, int KIND>
void run(const char *name, float *d)
{
    const int iters = 20000, grid = 148 * 3;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NP, NS, NA, KIND><<<grid, 256>>>(d, 100, 1.0001f, 1.f, 0x4B000000u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<NP, NS, NA, KIND><<<grid, 256>>>(d, iters, 1.0001f, 1.f, 0x4B000000u);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // 6 warps per scheduler share it: cycles per warp-iteration per scheduler = time * clock / (iters * 6)
    const double cyc = ms * 1e-3 * 1.965e9 / (iters * 6.0);
    printf("%-34s %8.3f ms  %7.2f cycles per warp-iteration   (instr: %d packed + %d scalar + %d alu; dispatch model %d)\n",
           name, ms, cyc, NP, NS, NA, 2 * NP + NS + NA);
}

int main()
{
    float *d; cudaMalloc(&d, 148 * 3 * 256 * 4);
    run<32, 0, 0, 0>("32 FFMA2", d);
    run<0, 32, 0, 0>("32 FFMA", d);
    run<0, 0, 32, 0>("32 PRMT", d);
    run<0, 0, 32, 1>("32 LOP3 (x2)", d);
    run<0, 0, 32, 2>("32 FMNMX", d);
    run<32, 0, 32, 0>("32 FFMA2 + 32 PRMT", d);
    run<32, 0, 16, 0>("32 FFMA2 + 16 PRMT", d);
    run<32, 0, 8, 0>("32 FFMA2 + 8 PRMT", d);
    run<16, 0, 32, 0>("16 FFMA2 + 32 PRMT", d);
    run<0, 32, 32, 0>("32 FFMA + 32 PRMT", d);
    run<16, 32, 0, 0>("16 FFMA2 + 32 FFMA", d);
    run<32, 0, 32, 2>("32 FFMA2 + 32 FMNMX", d);
    run<16, 16, 16, 0>("16 FFMA2 + 16 FFMA + 16 PRMT", d);
    return 0;
}
