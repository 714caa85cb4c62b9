// Microbenchmark: FP32 issue throughput on sm_100a, scalar vs packed (f32x2) add/mul/fma.
// Decides whether packing two channels per lane pays for the FIR inner loops, and documents that
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b)
{
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) { // scalar: 8 FMUL + 8 FADD on independent chains (no mul->add dependency)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __fmul_rn(x[i], a);
#pragma unroll
            for (int i = 8; i < 16; ++i) x[i] = __fadd_rn(x[i], b);
        } else if (MODE == 1) { // scalar FFMA x16
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], a, b);
        } else if (MODE == 2) { // packed FFMA2 x8 (16 lanes-values)
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float2 v = __ffma2_rn(make_float2(x[i], x[i + 1]), a2, b2);
                x[i] = v.x; x[i + 1] = v.y;
            }
        } else if (MODE == 3) { // packed: 4 FMUL2 + 4 FADD2 on independent chains
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float2 v = __fmul2_rn(make_float2(x[i], x[i + 1]), a2);
                x[i] = v.x; x[i + 1] = v.y;
            }
#pragma unroll
            for (int i = 8; i < 16; i += 2) {
                float2 v = __fadd2_rn(make_float2(x[i], x[i + 1]), b2);
                x[i] = v.x; x[i + 1] = v.y;
            }
        } else if (MODE == 4) { // scalar FMUL -> FADD dependent (the bit-exact FIR pattern), 16 values
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __fadd_rn(__fmul_rn(x[i], a), b);
        } else if (MODE == 5) { // packed FMUL2 then two scalar FADDs (unfusable bit-exact packed pattern)
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float2 v = __fmul2_rn(make_float2(x[i], x[i + 1]), a2);
                x[i] = __fadd_rn(v.x, b); x[i + 1] = __fadd_rn(v.y, b);
            }
        }
        else if (MODE == 6) { // FADD reg,reg (both operands registers, operands rotate)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __fadd_rn(x[i], x[(i + 5) & 15]);
        } else if (MODE == 7) { // FMUL reg,reg
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __fmul_rn(x[i], x[(i + 5) & 15]);
        } else if (MODE == 8) { // FFMA reg,reg,reg
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], x[(i + 5) & 15], x[(i + 11) & 15]);
        } else if (MODE == 9) { // the bit-exact FIR tap: v = a+b (reg,reg); acc_j += v*c_j (3 filters, coef constant)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float v = __fadd_rn(x[i], x[12 + ((i + 1) & 3)]);
                x[8 + i] = __fadd_rn(x[8 + i], __fmul_rn(v, a));
                x[12 + i] = __fadd_rn(x[12 + i], __fmul_rn(v, b));
            }
        } else if (MODE == 10) { // MODE 9 + one PRMT (alu pipe) per 5 FP ops
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float v = __fadd_rn(x[i], x[12 + ((i + 1) & 3)]);
                x[8 + i] = __fadd_rn(x[8 + i], __fmul_rn(v, a));
                x[12 + i] = __fadd_rn(x[12 + i], __fmul_rn(v, b));
                x[i] = __uint_as_float(__byte_perm(__float_as_uint(x[i]), 0x3f800000u, 0x7610 + it));
            }
        } else if (MODE == 11) { // packed tap: v2 = add2; p2 = mul2(v2, c2); scalar adds (no FFMA2 contraction possible)
#pragma unroll
            for (int i = 0; i < 4; i += 2) {
                const float2 v = __fadd2_rn(make_float2(x[i], x[i + 1]), make_float2(x[12 + ((i + 2) & 3)], x[13 + ((i + 2) & 3)]));
                const float2 pa = __fmul2_rn(v, a2), pb = __fmul2_rn(v, b2);
                x[8 + i] = __fadd_rn(x[8 + i], pa.x); x[9 + i] = __fadd_rn(x[9 + i], pa.y);
                x[12 + i] = __fadd_rn(x[12 + i], pb.x); x[13 + i] = __fadd_rn(x[13 + i], pb.y);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, double lane_ops_per_iter)
{
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    float *out;
    cudaMalloc(&out, blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<blocks, threads>>>(out, iters, 1.0001f, 0.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double thr_iters = 5.0 * blocks * threads * (double) iters;
    printf("%-44s %8.3f ms  %8.2f T f32-lane-ops/s\n", name, ms, thr_iters * lane_ops_per_iter / ms * 1e-9);
    cudaFree(out);
}

// numeric probe: does mul.rn.f32x2 -> add.rn.f32x2 round twice (like scalar .rn) or once (fused)?
__global__ void probe(const float *a, const float *b, const float *c, int n, int *n_like_fma, int *n_like_sep)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sep = __fadd_rn(__fmul_rn(a[i], b[i]), c[i]);
    const float fus = __fmaf_rn(a[i], b[i], c[i]);
    const float2 p = __fadd2_rn(__fmul2_rn(make_float2(a[i], a[i]), make_float2(b[i], b[i])), make_float2(c[i], c[i]));
    if (sep != fus) {
        if (p.x == fus) atomicAdd(n_like_fma, 1);
        if (p.x == sep) atomicAdd(n_like_sep, 1);
    }
}

int main()
{
    run<0>("scalar FMUL, FADD independent", 16);
    run<4>("scalar FMUL->FADD dependent", 32);
    run<1>("scalar FFMA (counted as 1 op/value)", 16);
    run<2>("packed FFMA2 (1 op/value)", 16);
    run<3>("packed FMUL2, FADD2 independent", 16);
    run<5>("packed FMUL2 -> 2x scalar FADD", 32);
    run<6>("scalar FADD reg,reg", 16);
    run<7>("scalar FMUL reg,reg", 16);
    run<8>("scalar FFMA reg,reg,reg", 16);
    run<9>("FIR tap scalar: 4x(FADD + 2 FMUL + 2 FADD)", 20);
    run<10>("FIR tap scalar + 4 PRMT (PRMT not counted)", 20);
    run<11>("FIR tap packed: FADD2 + 2 FMUL2 + 4 FADD (x2)", 20);
    const int n = 1 << 20;
    float *h = (float *) malloc(3 * n * 4);
    srand(1);
    for (int i = 0; i < 3 * n; ++i) h[i] = (float) rand() / RAND_MAX * 2.f - 1.f;
    float *d; int *cnt, hc[2];
    cudaMalloc(&d, 3 * n * 4); cudaMalloc(&cnt, 8); cudaMemset(cnt, 0, 8);
    cudaMemcpy(d, h, 3 * n * 4, cudaMemcpyHostToDevice);
    probe<<<n / 256, 256>>>(d, d + n, d + 2 * n, n, cnt, cnt + 1);
    cudaMemcpy(hc, cnt, 8, cudaMemcpyDeviceToHost);
    printf("probe: of the inputs where fused != separate: packed mul.rn+add.rn matched fused %d times, separate %d times\n", hc[0], hc[1]);
    return 0;
}
