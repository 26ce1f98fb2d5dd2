// microbench.cu — measures the machine limits the path-tracing kernels are judged against
// (SURVEY §8d asks for L2 bandwidth and FP32 issue rate next to the HBM roofline) and the cost
// of the primitives the wavefront queues are built from (same-address atomics, random 16/32/64-byte
// gathers, vector reductions).  Stand-alone: nvcc -O3 -gencode arch=compute_100a,code=sm_100a.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x)                                                                                                          \
    do {                                                                                                               \
        cudaError_t e = (x);                                                                                           \
        if (e != cudaSuccess) {                                                                                        \
            std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));                             \
            std::exit(1);                                                                                              \
        }                                                                                                              \
    } while (0)

template <class F> float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        best = ms < best ? ms : best;
    }
    return best;
}

// ---- atomics: one lane per warp does atomicAdd with return; `stride_words` spreads warps over counters
__global__ void k_atomic_ret(uint32_t *ctr, int n_ctr, int stride_words, int iters, uint32_t *sink) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t acc = 0;
    uint32_t *p = ctr + size_t(warp % n_ctr) * stride_words;
    for (int i = 0; i < iters; ++i) {
        uint32_t v = 0;
        if ((threadIdx.x & 31) == 0)
            v = atomicAdd(p, 32u);
        acc += __shfl_sync(0xffffffffu, v, 0);
    }
    if (acc == 0xdeadbeef)
        *sink = acc;
}
// the same without a return value (RED)
__global__ void k_atomic_red(uint32_t *ctr, int n_ctr, int stride_words, int iters) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t *p = ctr + size_t(warp % n_ctr) * stride_words;
    for (int i = 0; i < iters; ++i)
        if ((threadIdx.x & 31) == 0)
            atomicAdd(p, 32u);
}
// 7 lanes of a warp hit 7 counters of one 128-byte line in ONE instruction
__global__ void k_atomic_multi(uint32_t *ctr, int n_sets, int iters, uint32_t *sink) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    uint32_t acc = 0;
    uint32_t *p = ctr + size_t(warp % n_sets) * 32 + lane;
    for (int i = 0; i < iters; ++i) {
        uint32_t v = 0;
        if (lane < 7)
            v = atomicAdd(p, 3u);
        acc += __shfl_sync(0xffffffffu, v, i % 7);
    }
    if (acc == 0xdeadbeef)
        *sink = acc;
}

// ---- float4 reductions into a framebuffer at pseudo-random pixels
__global__ void k_red_v4(float4 *img, uint32_t npix, int iters) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        atomicAdd(img + (x % npix), make_float4(1.f, 2.f, 3.f, 0.f));
    }
}

// ---- random gathers of `VEC` x 16 bytes from a big array (path-state access pattern)
template <int VEC> __global__ void k_gather(const float4 *src, uint32_t n_rec, int iters, float *sink) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 777u;
    float acc = 0.f;
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        const float4 *p = src + size_t(x % n_rec) * VEC;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float4 a = __ldcs(p + v);
            acc += a.x + a.w;
        }
    }
    if (acc == 123.456f)
        *sink = acc;
}

// ---- streaming read (L2-resident when the buffer is small, HBM when large)
__global__ void k_stream(const float4 *src, size_t n, int passes, float *sink) {
    float acc = 0.f;
    for (int p = 0; p < passes; ++p)
        for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
            const float4 a = src[i];
            acc += a.x + a.y + a.z + a.w;
        }
    if (acc == 123.456f)
        *sink = acc;
}

// ---- FP32 FMA issue rate: 8 independent chains per thread
__global__ void k_fma(float *sink, int iters) {
    float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f, a4 = 4.f, a5 = 5.f, a6 = 6.f, a7 = 7.f;
    const float b = 1.0000001f, c = 0.5f;
    for (int i = 0; i < iters; ++i) {
        a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
        a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
    }
    const float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456f)
        *sink = s;
}

// ---- shared-memory bandwidth: LDS.128 from conflict-free addresses
__global__ void k_smem(float *sink, int iters) {
    __shared__ float4 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x)
        buf[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float acc = 0.f;
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
        const float4 a = buf[idx & 1023];
        acc += a.x + a.w;
        idx += 128;
    }
    if (acc == 123.456f)
        *sink = acc;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    std::printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d, \"l2_mb\": %.1f", prop.name, sms, clk_khz / 1000,
                prop.l2CacheSize / 1048576.0);
    float *sink;
    CK(cudaMalloc(&sink, 64));
    uint32_t *ctr;
    CK(cudaMalloc(&ctr, 1 << 22));
    CK(cudaMemset(ctr, 0, 1 << 22));
    const int grid = sms * 8, block = 128, warps = grid * block / 32;
    // atomics
    {
        const int iters = 64;
        struct { const char *name; int n_ctr, stride; } cases[] = {
            {"atomic_ret_1addr", 1, 32}, {"atomic_ret_32lines", 32, 32}, {"atomic_ret_128lines", 128, 32},
            {"atomic_ret_7addr_1line", 7, 1}, {"atomic_ret_4096lines", 4096, 32}};
        for (auto &c : cases) {
            float ms = time_ms([&] { k_atomic_ret<<<grid, block>>>(ctr, c.n_ctr, c.stride, iters, (uint32_t *)sink); });
            std::printf(", \"%s_gops\": %.3f", c.name, double(warps) * iters / ms / 1e6);
        }
        float ms = time_ms([&] { k_atomic_red<<<grid, block>>>(ctr, 1, 32, iters); });
        std::printf(", \"atomic_red_1addr_gops\": %.3f", double(warps) * iters / ms / 1e6);
        ms = time_ms([&] { k_atomic_multi<<<grid, block>>>(ctr, 1, iters, (uint32_t *)sink); });
        std::printf(", \"atomic_multi7_1line_ginstr\": %.3f", double(warps) * iters / ms / 1e6);
        ms = time_ms([&] { k_atomic_multi<<<grid, block>>>(ctr, 32, iters, (uint32_t *)sink); });
        std::printf(", \"atomic_multi7_32lines_ginstr\": %.3f", double(warps) * iters / ms / 1e6);
    }
    // float4 reductions
    {
        const uint32_t npix = 3840 * 2160;
        float4 *img;
        CK(cudaMalloc(&img, size_t(npix) * 16));
        CK(cudaMemset(img, 0, size_t(npix) * 16));
        const int iters = 64;
        float ms = time_ms([&] { k_red_v4<<<grid, block>>>(img, npix, iters); });
        std::printf(", \"red_v4_random_4k_gops\": %.3f", double(grid) * block * iters / ms / 1e6);
        ms = time_ms([&] { k_red_v4<<<grid, block>>>(img, 600 * 600, iters); });
        std::printf(", \"red_v4_random_600x600_gops\": %.3f", double(grid) * block * iters / ms / 1e6);
        CK(cudaFree(img));
    }
    // gathers + streams
    {
        const size_t bytes = size_t(1) << 30; // 1 GiB >> L2
        float4 *buf;
        CK(cudaMalloc(&buf, bytes));
        CK(cudaMemset(buf, 1, bytes));
        const int iters = 32;
        const int g2 = sms * 16;
        float ms = time_ms([&] { k_gather<1><<<g2, block>>>(buf, uint32_t(bytes / 16), iters, sink); });
        std::printf(", \"gather16_gbs\": %.1f", double(g2) * block * iters * 16 / ms / 1e6);
        ms = time_ms([&] { k_gather<2><<<g2, block>>>(buf, uint32_t(bytes / 32), iters, sink); });
        std::printf(", \"gather32_gbs\": %.1f", double(g2) * block * iters * 32 / ms / 1e6);
        ms = time_ms([&] { k_gather<4><<<g2, block>>>(buf, uint32_t(bytes / 64), iters, sink); });
        std::printf(", \"gather64_gbs\": %.1f", double(g2) * block * iters * 64 / ms / 1e6);
        ms = time_ms([&] { k_gather<8><<<g2, block>>>(buf, uint32_t(bytes / 128), iters, sink); });
        std::printf(", \"gather128_gbs\": %.1f", double(g2) * block * iters * 128 / ms / 1e6);
        ms = time_ms([&] { k_stream<<<sms * 16, 256>>>(buf, bytes / 16, 1, sink); });
        std::printf(", \"hbm_stream_read_gbs\": %.1f", double(bytes) / ms / 1e6);
        const size_t small = size_t(32) << 20; // 32 MiB: L2 resident
        ms = time_ms([&] { k_stream<<<sms * 16, 256>>>(buf, small / 16, 16, sink); });
        std::printf(", \"l2_stream_read_gbs\": %.1f", double(small) * 16 / ms / 1e6);
        // L2-resident gathers of 32 B (BVH node fetch pattern) from a 64 MiB table
        ms = time_ms([&] { k_gather<2><<<g2, block>>>(buf, uint32_t((size_t(64) << 20) / 32), iters * 4, sink); });
        std::printf(", \"l2_gather32_64mb_gbs\": %.1f", double(g2) * block * iters * 4 * 32 / ms / 1e6);
        ms = time_ms([&] { k_gather<2><<<g2, block>>>(buf, uint32_t((size_t(8) << 20) / 32), iters * 4, sink); });
        std::printf(", \"l2_gather32_8mb_gbs\": %.1f", double(g2) * block * iters * 4 * 32 / ms / 1e6);
        CK(cudaFree(buf));
    }
    // FP32 / shared memory
    {
        const int iters = 4096;
        float ms = time_ms([&] { k_fma<<<sms * 16, 256>>>(sink, iters); });
        std::printf(", \"fp32_fma_tflops\": %.2f", double(sms) * 16 * 256 * iters * 8 * 2 / ms / 1e9);
        ms = time_ms([&] { k_smem<<<sms * 8, 128>>>(sink, iters * 4); });
        std::printf(", \"smem_lds128_gbs\": %.1f", double(sms) * 8 * 128 * iters * 4 * 16 / ms / 1e6);
    }
    std::printf("}\n");
    return 0;
}
