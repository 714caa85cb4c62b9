// FMA-pipe cost of the packed f32x2 instructions the FIR loops are made of (sm_100a), measured with
// the real kernel's residency (3 CTAs x 256 threads per SM = 6 warps per scheduler).
// One "tap" = the work of fmb_demod_kernel's FIR1 per tap and thread: 4 pair sums, 12 products, 12 accumulations.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256, 3) k(float2 *out, int iters, float c0, float c1, float c2, float one, const float2 *in)
{
    float2 acc[12], w[8];
    const float2 one2 = make_float2(one, one);
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = in[threadIdx.x * 8 + i];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float2 m0 = make_float2(c0, c0), m1 = make_float2(c1, c1), m2 = make_float2(c2, c2);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (MODE == 0) { // the shipped loop: FADD2 + 3 FMUL2 + 3 FFMA2(one)
                    const float2 v = __fadd2_rn(w[(r + kk) & 3], w[4 + ((r - kk) & 3)]); w[(r + kk) & 3] = v; /* feedback: no two pair sums alike */
                    acc[r] = __ffma2_rn(__fmul2_rn(v, m0), one2, acc[r]);
                    acc[4 + r] = __ffma2_rn(__fmul2_rn(v, m1), one2, acc[4 + r]);
                    acc[8 + r] = __ffma2_rn(__fmul2_rn(v, m2), one2, acc[8 + r]);
                } else if (MODE == 1) { // previous loop: FADD2 + 3 FMUL2 + 6 scalar FADD
                    const float2 v = __fadd2_rn(w[(r + kk) & 3], w[4 + ((r - kk) & 3)]); w[(r + kk) & 3] = v; /* feedback: no two pair sums alike */
                    const float2 p0 = __fmul2_rn(v, m0), p1 = __fmul2_rn(v, m1), p2 = __fmul2_rn(v, m2);
                    acc[r].x = __fadd_rn(acc[r].x, p0.x); acc[r].y = __fadd_rn(acc[r].y, p0.y);
                    acc[4 + r].x = __fadd_rn(acc[4 + r].x, p1.x); acc[4 + r].y = __fadd_rn(acc[4 + r].y, p1.y);
                    acc[8 + r].x = __fadd_rn(acc[8 + r].x, p2.x); acc[8 + r].y = __fadd_rn(acc[8 + r].y, p2.y);
                } else if (MODE == 2) { // all scalar: 2 FADD + 6 FMUL + 6 FADD
                    const float2 a = w[(r + kk) & 3], b = w[4 + ((r - kk) & 3)];
                    const float vx = __fadd_rn(a.x, b.x), vy = __fadd_rn(a.y, b.y); w[(r + kk) & 3] = make_float2(vx, vy);
                    acc[r].x = __fadd_rn(acc[r].x, __fmul_rn(vx, c0)); acc[r].y = __fadd_rn(acc[r].y, __fmul_rn(vy, c0));
                    acc[4 + r].x = __fadd_rn(acc[4 + r].x, __fmul_rn(vx, c1)); acc[4 + r].y = __fadd_rn(acc[4 + r].y, __fmul_rn(vy, c1));
                    acc[8 + r].x = __fadd_rn(acc[8 + r].x, __fmul_rn(vx, c2)); acc[8 + r].y = __fadd_rn(acc[8 + r].y, __fmul_rn(vy, c2));
                } else if (MODE == 3) { // fused: FADD2 + 3 FFMA2 (FMB_PRECISION_FMA)
                    const float2 v = __fadd2_rn(w[(r + kk) & 3], w[4 + ((r - kk) & 3)]); w[(r + kk) & 3] = v; /* feedback: no two pair sums alike */
                    acc[r] = __ffma2_rn(v, m0, acc[r]); acc[4 + r] = __ffma2_rn(v, m1, acc[4 + r]); acc[8 + r] = __ffma2_rn(v, m2, acc[8 + r]);
                } else if (MODE == 4) { // 7 FFMA2(one) only
                    const float2 v = w[(r + kk) & 3];
                    acc[r] = __ffma2_rn(v, one2, acc[r]); acc[4 + r] = __ffma2_rn(v, one2, acc[4 + r]); acc[8 + r] = __ffma2_rn(v, one2, acc[8 + r]);
                    acc[r] = __ffma2_rn(v, one2, acc[r]); acc[4 + r] = __ffma2_rn(v, one2, acc[4 + r]); acc[8 + r] = __ffma2_rn(v, one2, acc[8 + r]);
                    acc[r] = __ffma2_rn(v, one2, acc[r]);
                } else if (MODE == 5) { // 7 FMUL2 only
                    float2 v = w[(r + kk) & 3];
                    acc[r] = __fmul2_rn(acc[r], m0); acc[4 + r] = __fmul2_rn(acc[4 + r], m1); acc[8 + r] = __fmul2_rn(acc[8 + r], m2);
                    acc[r] = __fmul2_rn(acc[r], m1); acc[4 + r] = __fmul2_rn(acc[4 + r], m2); acc[8 + r] = __fmul2_rn(acc[8 + r], m0);
                    acc[r] = __fmul2_rn(acc[r], v);
                } else if (MODE == 6) { // 7 FADD2 only
                    const float2 v = w[(r + kk) & 3];
                    acc[r] = __fadd2_rn(acc[r], v); acc[4 + r] = __fadd2_rn(acc[4 + r], v); acc[8 + r] = __fadd2_rn(acc[8 + r], v);
                    acc[r] = __fadd2_rn(acc[r], v); acc[4 + r] = __fadd2_rn(acc[4 + r], v); acc[8 + r] = __fadd2_rn(acc[8 + r], v);
                    acc[r] = __fadd2_rn(acc[r], v);
                } else if (MODE == 7) { // 14 scalar FADD
                    const float2 v = w[(r + kk) & 3];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        acc[r].x = __fadd_rn(acc[r].x, v.x); acc[r].y = __fadd_rn(acc[r].y, v.y);
                        acc[4 + r].x = __fadd_rn(acc[4 + r].x, v.x); acc[4 + r].y = __fadd_rn(acc[4 + r].y, v.y);
                        acc[8 + r].x = __fadd_rn(acc[8 + r].x, v.x); acc[8 + r].y = __fadd_rn(acc[8 + r].y, v.y);
                    }
                    acc[r].x = __fadd_rn(acc[r].x, v.y); acc[r].y = __fadd_rn(acc[r].y, v.x);
                } else if (MODE == 8) { // 14 scalar FFMA (3 registers)
                    const float2 v = w[(r + kk) & 3];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        acc[r].x = __fmaf_rn(v.x, c0, acc[r].x); acc[r].y = __fmaf_rn(v.y, c0, acc[r].y);
                        acc[4 + r].x = __fmaf_rn(v.x, c1, acc[4 + r].x); acc[4 + r].y = __fmaf_rn(v.y, c1, acc[4 + r].y);
                        acc[8 + r].x = __fmaf_rn(v.x, c2, acc[8 + r].x); acc[8 + r].y = __fmaf_rn(v.y, c2, acc[8 + r].y);
                    }
                    acc[r].x = __fmaf_rn(v.y, c0, acc[r].x); acc[r].y = __fmaf_rn(v.x, c0, acc[r].y);
                }
            }
            w[kk & 3].x += 1.0f; // keep the windows changing (1 scalar FADD per tap)
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 12; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int lane_ops)
{
    const int blocks = 148 * 3, threads = 256, iters = 2048;
    float2 *out, *in;
    cudaMalloc(&out, blocks * threads * 8); cudaMalloc(&in, 256 * 8 * 8); cudaMemset(in, 0, 256 * 8 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0.9999f, 1.0002f, 1.0f, in);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0.9999f, 1.0002f, 1.0f, in);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // 6 warps per scheduler, 4 taps per iteration
    const double cyc = ms * 1e-3 / 5 * 1.965e9 / iters / 4 / 6.0;
    printf("%-44s %8.3f ms  %6.2f scheduler cycles per warp-tap; %3d lane-ops -> %5.2f lane-op/clk/lane\n", name, ms / 5, cyc, lane_ops, lane_ops / cyc);
    cudaFree(out); cudaFree(in);
}

int main()
{
    run<0>("4 FADD2 + 12 FMUL2 + 12 FFMA2(one)", 56);
    run<1>("4 FADD2 + 12 FMUL2 + 24 FADD", 56);
    run<2>("8 FADD + 24 FMUL + 24 FADD", 56);
    run<3>("4 FADD2 + 12 FFMA2 (fused)", 32);
    run<4>("28 FFMA2(one)", 56);
    run<5>("28 FMUL2", 56);
    run<6>("28 FADD2", 56);
    run<7>("56 FADD", 56);
    run<8>("56 FFMA", 56);
    return 0;
}
