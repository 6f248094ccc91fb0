// membw.cu — hardware numbers that bound the phase-space kernels (run on the B200 box):
//   1. streaming read bandwidth from DRAM with per-thread LDG.64 / LDG.128
//   2. the same loop over an L2-resident buffer (L2 -> SM bandwidth: the gather kernels read every plane ~5x through L2)
//   3. cp.async.bulk (TMA 1-D bulk copy, global -> shared) streaming, DRAM and L2 resident
//   4. fp64 FMA issue rate
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membw tools/membw.cu ; run: ./membw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int VEC>
__global__ void __launch_bounds__(256) read_kernel(const double* __restrict__ p, size_t n, int reps, double* out) {
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x * VEC;
    for (int r = 0; r < reps; ++r) {
        for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i + VEC <= n; i += stride) {
            if (VEC == 2) {
                const double2 v = *reinterpret_cast<const double2*>(p + i);
                acc += v.x + v.y;
            } else {
                acc += p[i];
            }
        }
    }
    if (acc == 123.456) out[0] = acc;
}

// bulk copies: each CTA streams chunks of CH bytes through a ring of STAGES shared-memory buffers; one thread issues,
// all threads touch one value per chunk so the data is consumed
template <int CH, int STAGES>
__global__ void __launch_bounds__(128) bulk_kernel(const double* __restrict__ p, size_t nchunks, int reps, double* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar[STAGES];
    double* buf = reinterpret_cast<double*>(smem);
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(bar);
    const unsigned buf0 = (unsigned)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    double acc = 0.0;
    size_t total = 0;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) ++total;
    total *= reps;
    auto src_of = [&](size_t k) { return p + ((blockIdx.x + (k % (total / reps)) * gridDim.x) * (size_t)(CH / 8)); };
    auto issue = [&](size_t k) {
        const int s = k % STAGES;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * s), "r"(CH) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         buf0 + s * CH),
                     "l"(src_of(k)), "r"(CH), "r"(bar0 + 8 * s)
                     : "memory");
    };
    if (threadIdx.x == 0)
        for (size_t k = 0; k < STAGES && k < total; ++k) issue(k);
    for (size_t k = 0; k < total; ++k) {
        const int s = k % STAGES;
        const unsigned parity = (k / STAGES) & 1;
        unsigned done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar0 + 8 * s), "r"(parity) : "memory");
        }
        acc += buf[s * (CH / 8) + threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0 && k + STAGES < total) issue(k + STAGES);
    }
    if (acc == 123.456) out[0] = acc;
}

__global__ void __launch_bounds__(256) fma_kernel(double* out, int iters) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double x = 1.0000001, y = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
        a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <class F>
float time_ms(F f, int n = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < n; ++i) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    const size_t big = (size_t)4 << 30;       // 4 GiB: DRAM
    const size_t small = (size_t)48 << 20;    // 48 MiB: L2 resident
    double *p, *out;
    CK(cudaMalloc(&p, big));
    CK(cudaMalloc(&out, 1 << 22));
    CK(cudaMemset(p, 0, big));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
    const int grids[] = {148 * 4, 148 * 8, 148 * 16};
    for (int g : grids) {
        float t1 = time_ms([&] { read_kernel<1><<<g, 256>>>(p, big / 8, 1, out); });
        float t2 = time_ms([&] { read_kernel<2><<<g, 256>>>(p, big / 8, 1, out); });
        printf("DRAM read  grid %5d: LDG.64 %7.1f GB/s   LDG.128 %7.1f GB/s\n", g, big / t1 / 1e6, big / t2 / 1e6);
        const int reps = 64;
        float t3 = time_ms([&] { read_kernel<1><<<g, 256>>>(p, small / 8, reps, out); });
        float t4 = time_ms([&] { read_kernel<2><<<g, 256>>>(p, small / 8, reps, out); });
        printf("L2   read  grid %5d: LDG.64 %7.1f GB/s   LDG.128 %7.1f GB/s\n", g, small * (double)reps / t3 / 1e6,
               small * (double)reps / t4 / 1e6);
    }
    {
        constexpr int CH = 16384, ST = 4;
        auto k = bulk_kernel<CH, ST>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * ST));
        for (int g : {148, 148 * 2, 148 * 3}) {
            float t1 = time_ms([&] { k<<<g, 128, CH * ST>>>(p, big / CH, 1, out); });
            float t2 = time_ms([&] { k<<<g, 128, CH * ST>>>(p, small / CH, 64, out); });
            printf("bulk copy 16 KB x4 stages, grid %4d: DRAM %7.1f GB/s   L2 %7.1f GB/s\n", g, big / t1 / 1e6,
                   small * 64.0 / t2 / 1e6);
        }
    }
    {
        constexpr int CH = 4096, ST = 8;
        auto k = bulk_kernel<CH, ST>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * ST));
        for (int g : {148 * 2, 148 * 4}) {
            float t1 = time_ms([&] { k<<<g, 128, CH * ST>>>(p, big / CH, 1, out); });
            float t2 = time_ms([&] { k<<<g, 128, CH * ST>>>(p, small / CH, 64, out); });
            printf("bulk copy  4 KB x8 stages, grid %4d: DRAM %7.1f GB/s   L2 %7.1f GB/s\n", g, big / t1 / 1e6,
                   small * 64.0 / t2 / 1e6);
        }
    }
    {
        const int iters = 1 << 16;
        for (int g : {148 * 4, 148 * 8}) {
            float t = time_ms([&] { fma_kernel<<<g, 256>>>(out, iters); });
            printf("fp64 FMA grid %5d: %.2f TFLOP/s (%.1f DFMA/clk/SM at 1.965 GHz)\n", g,
                   2.0 * 8 * iters * g * 256.0 / t / 1e9, 8.0 * iters * g * 256.0 / (t * 1e-3) / 148 / 1.965e9);
        }
    }
    CK(cudaGetLastError());
    return 0;
}
