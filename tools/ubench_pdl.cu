// ubench_pdl.cu -- does programmatic dependent launch hide the ramp/drain between back-to-back full-wave kernels?
//   A: the stand-in for the demod kernel: 444 CTAs (3 per SM: 72 KB smem, 256 threads, 80 regs), ragged work per CTA,
//      griddepcontrol.launch_dependents at its start, per-CTA done flags; CTA i of launch b+1 first waits for CTA i of launch b
//   D: the stand-in for a co-resident de-emphasis kernel: 148 CTAs x 128 threads, <= 32 regs, 8 KB smem; waits until all
//      CTAs of its A are done
// modes: 0 plain stream order A,A,..   1 A[PDL] chain   2 A[PDL],D[PDL] chain   3 mode 1 + cudaEventRecord between launches
//        4 mode 2 without the PDL attribute (plain A,D,A,D)
// prints ms per step and, per launch, when the first CTA started / the last CTA ended (us, relative to the first launch).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

struct Rec { unsigned long long first_start, last_end; };

__global__ void __launch_bounds__(256, 3) kernA(int step, unsigned *done, Rec *rec, int unit_ns, int use_pdl, unsigned *err)
{
    extern __shared__ float sm[];
    if (use_pdl) asm volatile("griddepcontrol.launch_dependents;");
    const unsigned long long t0 = gtime();
    if (threadIdx.x == 0) {
        atomicMin(&rec[step].first_start, t0);
        if (step > 0) {                           // dependency on the same CTA of the previous launch (flag = step)
            volatile unsigned *f = done + blockIdx.x;
            unsigned long long spins = 0;
            while (*f < (unsigned) step) { __nanosleep(200); if (++spins > 20000000ull) { atomicAdd(err, 1u); break; } }
            __threadfence();
        }
    }
    __syncthreads();
    // ragged work: 18 or 19 units of unit_ns
    const int units = 18 + ((blockIdx.x * 7 + step) % 20 < 9 ? 1 : 0);
    const unsigned long long t_end = gtime() + (unsigned long long) units * unit_ns;
    float a = threadIdx.x, b = 1.0001f;
    while (gtime() < t_end) {
#pragma unroll
        for (int i = 0; i < 64; ++i) a = a * b + 0.5f;
    }
    sm[threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicExch(done + blockIdx.x, (unsigned) step + 1u);
        atomicMax(&rec[step].last_end, gtime());
    }
}

__global__ void __launch_bounds__(128) kernD(int step, const unsigned *done, int n_a, Rec *rec, int use_pdl, unsigned *err)
{
    __shared__ float buf[2048];
    if (use_pdl) asm volatile("griddepcontrol.launch_dependents;");
    if (threadIdx.x == 0) atomicMin(&rec[step].first_start, gtime());
    // wait for "my" share of A's CTAs
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_a; i += gridDim.x * blockDim.x) {
        volatile const unsigned *f = done + i;
        unsigned long long spins = 0;
        while (*f < (unsigned) step + 1u) { __nanosleep(500); if (++spins > 8000000ull) { atomicAdd(err, 1u); break; } }
    }
    buf[threadIdx.x] = 1.f;
    __syncthreads();
    if (threadIdx.x == 0) atomicMax(&rec[step].last_end, gtime());
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

static void launch(void (*k)(int, unsigned *, Rec *, int, int, unsigned *), int grid, int block, size_t smem, cudaStream_t s, bool pdl,
                   int step, unsigned *done, Rec *rec, int unit_ns, unsigned *err)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    int up = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, k, step, done, rec, unit_ns, up, err));
}
static void launchD(int grid, cudaStream_t s, bool pdl, int step, const unsigned *done, int n_a, Rec *rec, unsigned *err)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    int up = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kernD, step, done, n_a, rec, up, err));
}

int main(int argc, char **argv)
{
    const int unit_ns = argc > 1 ? atoi(argv[1]) : 19000, steps = 12;
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t smem = 72 * 1024;
    CK(cudaFuncSetAttribute(kernA, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernA, 256, smem));
    cudaFuncAttributes fa, fd; CK(cudaFuncGetAttributes(&fa, kernA)); CK(cudaFuncGetAttributes(&fd, kernD));
    const int grid = occ * sms;
    printf("SMs %d, A: %d regs, occupancy %d -> grid %d; D: %d regs; unit %d ns\n", sms, fa.numRegs, occ, grid, fd.numRegs, unit_ns);
    unsigned *done, *err; Rec *rec;
    CK(cudaMalloc(&done, grid * sizeof(unsigned))); CK(cudaMalloc(&err, 4)); CK(cudaMalloc(&rec, steps * 2 * sizeof(Rec)));
    cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, ev; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (int mode = 0; mode < 5; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaMemset(done, 0, grid * sizeof(unsigned))); CK(cudaMemset(err, 0, 4));
            std::vector<Rec> init(steps * 2, Rec{~0ull, 0ull});
            CK(cudaMemcpy(rec, init.data(), steps * 2 * sizeof(Rec), cudaMemcpyHostToDevice));
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, s));
            for (int b = 0; b < steps; ++b) {
                const bool pdl = (mode == 1 || mode == 2 || mode == 3);
                launch(kernA, grid, 256, smem, s, pdl, b, done, rec, unit_ns, err);
                if (mode == 3) CK(cudaEventRecord(ev, s));
                if (mode == 2 || mode == 4) launchD(sms, s, pdl, b, done, grid, rec + steps, err);
            }
            CK(cudaEventRecord(e1, s));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            std::vector<Rec> r(steps * 2); unsigned nerr = 0;
            CK(cudaMemcpy(r.data(), rec, steps * 2 * sizeof(Rec), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&nerr, err, 4, cudaMemcpyDeviceToHost));
            if (rep == 0) continue;
            printf("mode %d: %.4f ms per step (%d steps), spin timeouts %u\n", mode, ms / steps, steps, nerr);
            const unsigned long long t0 = r[0].first_start;
            for (int b = 0; b < 4; ++b) {
                printf("   A(%d): first CTA start %8.1f us, last CTA end %8.1f us", b, (r[b].first_start - t0) * 1e-3, (r[b].last_end - t0) * 1e-3);
                if (mode == 2 || mode == 4) printf("   D(%d): first start %8.1f, last end %8.1f", b, (r[steps + b].first_start - t0) * 1e-3, (r[steps + b].last_end - t0) * 1e-3);
                printf("\n");
            }
        }
    }
    return 0;
}
