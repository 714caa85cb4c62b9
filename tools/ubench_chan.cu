// ubench_chan.cu -- the channel stage of fmb_demod_kernel in isolation: how many scheduler cycles does one channel-FIR
// output (both components) cost a warp, with and without the discriminator, at 2 / 4 / 6 warps per scheduler?
// Compare with the pipe bounds of the stage: FMA 47 packed x 2 = 94 cycles (+ 16 scalar with the discriminator),
// ALU (32 PRMT + 8 LOP3) x 2 = 80 cycles (+ ~20 x 2 with the discriminator).
#include "../rtl_fm_player_b200/csrc/fmb_kernels.cu"
#include <cstdio>

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
    return 0;
}
