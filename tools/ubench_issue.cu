// Does a packed f32x2 instruction (FMUL2/FADD2) free issue slots for other pipes on sm_100a?
// Per iteration and thread: NF FP32 lane-operations (scalar or packed) + NA LOP3 (alu pipe), all independent chains.
// Prints SM cycles per warp-iteration per scheduler; compare with the issue-slot and FMA-pipe counts.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, unsigned m, long long *cyc)
{
    float x[16]; unsigned y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = threadIdx.x * 0.001f + i; y[i] = threadIdx.x * 2654435761u + i; }
    const float2 a2 = make_float2(a, a);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 1 || MODE == 5) { // 16 scalar FMUL
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __fmul_rn(x[i], a);
        }
        if (MODE == 2 || MODE == 3 || MODE == 4) { // 8 FMUL2 = 16 lane-ops
#pragma unroll
            for (int i = 0; i < 16; i += 2) { float2 v = __fmul2_rn(make_float2(x[i], x[i + 1]), a2); x[i] = v.x; x[i + 1] = v.y; }
        }
        if (MODE == 1 || MODE == 3 || MODE == 6) { // 8 LOP3
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = y[i] ^ (y[(i + 1) & 7] & m);
        }
        if (MODE == 4 || MODE == 5 || MODE == 7) { // 16 LOP3
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] = y[i] ^ (y[(i + 1) & 15] & m);
        }
    }
    const long long t1 = clock64();
    float s = 0; unsigned u = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) { s += x[i]; u ^= y[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float) u;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char *name)
{
    const int blocks = 148 * 8, threads = 256, iters = 4096; // 64 warps per SM = 16 per scheduler
    float *out; long long *cyc;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0x55555555u, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0x55555555u, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // cycles per (warp, iteration) per scheduler at 1.965 GHz: 16 warps per scheduler
    const double cyc_per = ms * 1e-3 / 5 * 1.965e9 / iters / 16.0;
    printf("%-28s %8.3f ms   %6.2f scheduler cycles per warp-iteration (at 1965 MHz)\n", name, ms / 5, cyc_per);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("16 FMUL");
    run<2>("8 FMUL2");
    run<6>("8 LOP3");
    run<7>("16 LOP3");
    run<1>("16 FMUL + 8 LOP3");
    run<3>("8 FMUL2 + 8 LOP3");
    run<5>("16 FMUL + 16 LOP3");
    run<4>("8 FMUL2 + 16 LOP3");
    return 0;
}
