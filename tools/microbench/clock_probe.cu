// SM clock timeline probe (tuning tool, not part of the product library): one warp on one SM samples
// (globaltimer, clock64) every `interval_ns` while other kernels run on other streams; f(t) = d clock / d time.
// nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o tools/microbench/libclockprobe.so tools/microbench/clock_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void clock_probe_kernel(unsigned long long* out, int n, unsigned long long interval_ns) {
    if (threadIdx.x != 0) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned long long next = t0;
    for (int i = 0; i < n; ++i) {
        unsigned long long t;
        do {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t < next) __nanosleep(2000);
        } while (t < next);
        out[2 * i] = t;
        out[2 * i + 1] = clock64();
        next += interval_ns;
    }
}
extern "C" int clock_probe_launch(unsigned long long* out, int n, unsigned long long interval_ns, cudaStream_t s) {
    clock_probe_kernel<<<1, 32, 0, s>>>(out, n, interval_ns);
    return (int)cudaGetLastError();
}
