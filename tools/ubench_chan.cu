// ubench_chan.cu -- the channel stage of fmb_demod_kernel in isolation: how many scheduler cycles does one channel-FIR
// output (both components) cost a warp, with and without the discriminator, at 2 / 4 / 6 warps per scheduler?
// Compare with the pipe bounds of the stage: FMA 47 packed x 2 = 94 cycles (+ 16 scalar with the discriminator),
// ALU (32 PRMT + 8 LOP3) x 2 = 80 cycles (+ ~20 x 2 with the discriminator).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -prec-div=true -ftz=false -Iinclude -Irtl_fm_player_b200/csrc -o tools/ubench_chan tools/ubench_chan.cu rtl_fm_player_b200/csrc/build/fm_design.o
#include "../rtl_fm_player_b200/csrc/fmb_kernels.cu"
#include <cstdio>

template <bool ROT, bool FMA, typename Emit>
__device__ __forceinline__ void chan_fir_once(const unsigned rbase, const float *cs, const float2 one2, Emit emit)
{
    auto row = [&](const int j) { return lds128(rbase + (j >> 3) * RAW_PITCH + (j & 7) * 16); };
    float2 W[4][8];
    magic_row2<ROT, true>(row(1), W[1]);
    magic_row2<ROT, false>(row(2), W[2]);
    magic_row2<ROT, true>(row(3), W[3]);
#pragma unroll
    for (int o = 1; o < 9; ++o) {
        if ((o + 3) & 1) magic_row2<ROT, true>(row(o + 3), W[(o + 3) & 3]);
        else magic_row2<ROT, false>(row(o + 3), W[(o + 3) & 3]);
        const float2 (&R0)[8] = W[o & 3], (&R1)[8] = W[(o + 1) & 3], (&R2)[8] = W[(o + 2) & 3], (&R3)[8] = W[(o + 3) & 3];
        float2 acc = __fmul2_rn(__fadd2_rn(R0[0], R3[7]), make_float2(cs[0], cs[0]));
#pragma unroll
        for (int t = 1; t < 8; ++t) acc = mac2<FMA>(__fadd2_rn(R0[t], R3[7 - t]), cs[t], one2, acc);
#pragma unroll
        for (int t = 8; t < 16; ++t) acc = mac2<FMA>(__fadd2_rn(R1[t - 8], R2[15 - t]), cs[t], one2, acc);
        emit(o, acc.x, acc.y);
    }
}

// variants of the stage, all at 2 CTAs per SM (up to 128 registers): V 0 as built, 1 convert-once, 2 as built without
// rotation (half the complements), 3 as built + discriminator with the branch-free division, 4 convert-once + that
template <int V>
__global__ void __launch_bounds__(256, 2) kvar(float *out, int iters, const __grid_constant__ fmb_tables c)
{
    extern __shared__ __align__(16) unsigned char raw[];
    const int tid = threadIdx.x;
    for (int i = tid; i < RAW_BYTES / 4; i += 256) reinterpret_cast<unsigned *>(raw)[i] = i * 2654435761u + tid;
    __syncthreads();
    const float2 one2 = make_float2(c.one, c.one);
    const unsigned rbase = smem_addr(raw + tid * RAW_PITCH);
    float accx = 0.f, accy = 0.f;
    for (int it = 0; it < iters; ++it) {
        float pr = 0.f, pj = 0.f;
        auto emit = [&](const int o, float ai, float aq) {
            if (V >= 3) {
                const float y = sub(mul(pr, aq), mul(pj, ai));
                const float x = add(mul(ai, pr), mul(aq, pj));
                const bool steep = fabsf(x) < fabsf(y);
                accx += octant_finish(y, x, div_core(steep ? x : y, steep ? y : x));
            } else { accx += ai; accy += aq; }
            pr = ai; pj = aq;
        };
        if (V == 0 || V == 3) chan_fir_packed<true, false>(rbase, c.chan_s, one2, emit);
        else if (V == 2) chan_fir_packed<false, false>(rbase, c.chan_s, one2, emit);
        else chan_fir_once<true, false>(rbase, c.chan_s, one2, emit);
        __syncwarp();
    }
    out[blockIdx.x * 256 + tid] = accx + accy;
}

template <int V>
void runvar(const char *name, const fmb_tables &t, float *d)
{
    const int iters = 2000, smem = 100 * 1024;
    cudaFuncSetAttribute(kvar<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kvar<V>, 256, smem);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kvar<V>);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kvar<V><<<148 * occ, 256, smem>>>(d, 10, t);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kvar<V><<<148 * occ, 256, smem>>>(d, iters, t);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-44s %d CTAs/SM, %3d regs, %zu B local  %6.1f scheduler cycles per output-warp\n", name, occ, fa.numRegs, (size_t) fa.localSizeBytes,
           ms * 1e-3 * 1.965e9 / (iters * 8.0 * occ * 2));
}

template <int VARIANT>      // 0: FIR only, 1: FIR + discriminator (fdiv), 2: FIR + conversions only (no FP: emit the converted rows)
__global__ void __launch_bounds__(256, 3) kchan(float *out, int iters, const __grid_constant__ fmb_tables c)
{
    extern __shared__ __align__(16) unsigned char raw[];
    const int tid = threadIdx.x;
    for (int i = tid; i < RAW_BYTES / 4; i += 256) reinterpret_cast<unsigned *>(raw)[i] = i * 2654435761u + tid;
    __syncthreads();
    const float2 one2 = make_float2(c.one, c.one);
    const unsigned rbase = smem_addr(raw + tid * RAW_PITCH);
    float accx = 0.f, accy = 0.f;
    for (int it = 0; it < iters; ++it) {
        float pr = 0.f, pj = 0.f;
        chan_fir_packed<true, false>(rbase + (it & 1) * 0, c.chan_s, one2, [&](const int o, float ai, float aq) {
            if (VARIANT == 1) {
                const float y = sub(mul(pr, aq), mul(pj, ai));
                const float x = add(mul(ai, pr), mul(aq, pj));
                accx += octant_angle(y, x);
            } else { accx += ai; accy += aq; }
            pr = ai; pj = aq;
        });
        __syncwarp();
    }
    out[blockIdx.x * 256 + tid] = accx + accy;
}

template <int VARIANT>
void run(const char *name, int ctas_per_sm, const fmb_tables &t, float *d)
{
    const int iters = 2000;
    // occupancy is set by the dynamic shared memory size
    const int smem = ctas_per_sm == 3 ? 72 * 1024 : ctas_per_sm == 2 ? 100 * 1024 : 200 * 1024;
    cudaFuncSetAttribute(kchan<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kchan<VARIANT>, 256, smem);
    const int grid = 148 * occ;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kchan<VARIANT><<<grid, 256, smem>>>(d, 10, t);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kchan<VARIANT><<<grid, 256, smem>>>(d, iters, t);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // per scheduler: occ*2 warps, each iters*8 outputs
    const double cyc = ms * 1e-3 * 1.965e9 / (iters * 8.0 * occ * 2);
    printf("%-28s %d CTAs/SM (%d warps/scheduler)  %7.3f ms  %6.1f scheduler cycles per output-warp\n", name, occ, occ * 2, ms, cyc);
}

int main()
{
    fmb_config cfg; memset(&cfg, 0, sizeof cfg);
    cfg.rate_in = 192000; cfg.rate_out2 = 48000; cfg.mode = 2; cfg.size = 90; cfg.volume = 0.4f; cfg.deemph = 50e-6;
    fmb_tables t; fmb_design_tables(&cfg, &t);
    float *d; cudaMalloc(&d, 148 * 3 * 256 * 4);
    for (int c = 1; c <= 3; ++c) run<0>("channel FIR only", c, t, d);
    for (int c = 1; c <= 3; ++c) run<1>("channel FIR + discriminator", c, t, d);
    runvar<0>("FIR as built", t, d);
    runvar<1>("FIR, rows converted once", t, d);
    runvar<2>("FIR as built, no rotation", t, d);
    runvar<3>("FIR as built + branch-free discriminator", t, d);
    runvar<4>("FIR converted once + branch-free discr.", t, d);
    return 0;
}
