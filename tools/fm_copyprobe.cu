/*
 * fm_copyprobe.cu -- raw host<->device copy ceiling of the box (MEASUREMENT TOOL, not product code).
 *
 * End to end the demodulator is bound by the host->GPU copy fabric (2 bytes per IQ sample in, 0.125 out).
 * This probe measures what that fabric carries with nothing else going on: per device one pinned host buffer
 * (portable; optionally write-combined) and one device buffer, `reps` back-to-back cudaMemcpyAsync on one
 * stream per device, all devices at once, timed from the first enqueue to the last completion.
 *
 *   dir 0: H2D only     dir 1: D2H only     dir 2: H2D of `bytes` and D2H of `bytes/16` together (the
 *                                            demodulator's own ratio, 2 B in : 0.125 B out per IQ sample)
 *
 * Built as tools/libfmprobe.so (bench.py's e2e.copy_ceiling, tools/h2d_ceiling.py); plain C ABI.
 */
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {
struct Dev {
    int dev = 0;
    void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;
};
}

#define CKP(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { snprintf(err, 256, "%s: %s", #call, cudaGetErrorString(e_)); rc = -1; goto done; } } while (0)

/* `sync` (may be NULL) is called after the buffers are set up and the warm-up copies are done, right before the
 * timed copies start: a multi-process driver passes a barrier here so that all processes copy AT THE SAME TIME
 * (allocating and touching 256 MiB of pinned memory takes far longer than copying it).  t_begin/t_end (may be NULL)
 * receive the CLOCK_REALTIME seconds of the timed part, so that the driver can compute the aggregate rate over the
 * union of the processes' intervals. */
extern "C" int fmprobe_copy_sync(const int *devices, int n_dev, size_t bytes, int reps, int dir, int write_combined,
                                 double *gbs_h2d, double *gbs_d2h, double *seconds, char *err /* >= 256 bytes */,
                                 void (*sync)(void), double *t_begin, double *t_end)
{
    std::vector<Dev> dv((size_t) n_dev);
    int rc = 0;
    const size_t out_bytes = dir == 2 ? bytes / 16 : bytes;
    const bool do_in = dir != 1, do_out = dir != 0;
    double dt = 0;
    if (err) err[0] = 0;
    for (int i = 0; i < n_dev; ++i) {
        Dev &d = dv[(size_t) i];
        d.dev = devices[i];
        CKP(cudaSetDevice(d.dev));
        if (do_in) {
            CKP(cudaHostAlloc(&d.h_in, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
            memset(d.h_in, 0x5a, bytes);
            CKP(cudaMalloc(&d.d_in, bytes));
            CKP(cudaStreamCreateWithFlags(&d.s_in, cudaStreamNonBlocking));
        }
        if (do_out) {
            CKP(cudaHostAlloc(&d.h_out, out_bytes, cudaHostAllocPortable));
            CKP(cudaMalloc(&d.d_out, out_bytes));
            CKP(cudaMemset(d.d_out, 1, out_bytes));
            CKP(cudaStreamCreateWithFlags(&d.s_out, cudaStreamNonBlocking));
        }
    }
    for (int pass = 0; pass < 2; ++pass) {               /* pass 0: warm-up (2 copies), pass 1: timed */
        const int n = pass == 0 ? 2 : reps;
        for (int i = 0; i < n_dev; ++i) { CKP(cudaSetDevice(dv[(size_t) i].dev)); CKP(cudaDeviceSynchronize()); }
        if (pass == 1 && sync) sync();
        if (pass == 1 && t_begin) *t_begin = std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < n; ++r)
            for (int i = 0; i < n_dev; ++i) {
                Dev &d = dv[(size_t) i];
                CKP(cudaSetDevice(d.dev));
                if (do_in) CKP(cudaMemcpyAsync(d.d_in, d.h_in, bytes, cudaMemcpyHostToDevice, d.s_in));
                if (do_out) CKP(cudaMemcpyAsync(d.h_out, d.d_out, out_bytes, cudaMemcpyDeviceToHost, d.s_out));
            }
        for (int i = 0; i < n_dev; ++i) { CKP(cudaSetDevice(dv[(size_t) i].dev)); CKP(cudaDeviceSynchronize()); }
        dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (pass == 1 && t_end) *t_end = std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
    }
    if (seconds) *seconds = dt;
    if (gbs_h2d) *gbs_h2d = do_in ? (double) bytes * reps * n_dev / dt * 1e-9 : 0.0;
    if (gbs_d2h) *gbs_d2h = do_out ? (double) out_bytes * reps * n_dev / dt * 1e-9 : 0.0;
done:
    for (auto &d : dv) {
        cudaSetDevice(d.dev);
        if (d.h_in) cudaFreeHost(d.h_in);
        if (d.h_out) cudaFreeHost(d.h_out);
        if (d.d_in) cudaFree(d.d_in);
        if (d.d_out) cudaFree(d.d_out);
        if (d.s_in) cudaStreamDestroy(d.s_in);
        if (d.s_out) cudaStreamDestroy(d.s_out);
    }
    return rc;
}

extern "C" int fmprobe_copy(const int *devices, int n_dev, size_t bytes, int reps, int dir, int write_combined,
                            double *gbs_h2d, double *gbs_d2h, double *seconds, char *err)
{
    return fmprobe_copy_sync(devices, n_dev, bytes, reps, dir, write_combined, gbs_h2d, gbs_d2h, seconds, err, nullptr, nullptr, nullptr);
}
