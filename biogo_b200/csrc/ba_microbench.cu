// ba_microbench.cu -- measures the integer/DPX issue rates the roofline is quoted against.
//
// the DPX issue rate per SM is not documented, so the first GPU task
// is a micro-benchmark of independent DPX chains on all SMs.  Each kernel keeps NCH
// independent dependency chains per thread busy with one instruction kind, 32 warps per SM,
// and reports thread-level results per clock per SM (cycles from clock64, wall time from
// CUDA events).  Output: one JSON object per line on stdout.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define NCH 16
#define ITERS 8192

enum Op {
    OP_VIMAX3_S16X2, OP_VIADDMAX_S16X2_RELU, OP_VIADD_16X2, OP_VIMAX3_S32, OP_VIADDMAX_S32, OP_VIMIN3_U32,
    OP_IADD3, OP_LOP3, OP_PRMT, OP_IMAD, OP_MIX_DPX_IMAD, OP_HSET2, OP_SHFL, OP_LDS, OP_MIX_DPX_LOP, OP_VIMNMX_PRED,
    OP_COUNT
};
static const char *op_names[OP_COUNT] = {
    "vimax3_s16x2", "viaddmax_s16x2_relu", "viadd_16x2", "vimax3_s32", "viaddmax_s32", "vimin3_u32",
    "iadd3", "lop3", "prmt", "imad", "mix_vimax3_s16x2+imad", "hset2_eq", "shfl_up", "lds32", "mix_vimax3_s16x2+lop3",
    "vibmax_s16x2_pred"};

template <int OP>
__device__ __forceinline__ unsigned step(unsigned x, unsigned a, unsigned b, const unsigned *sm, int lane)
{
    if (OP == OP_VIMAX3_S16X2) return __vimax3_s16x2(x, a, b);
    if (OP == OP_VIADDMAX_S16X2_RELU) return __viaddmax_s16x2_relu(x, a, b);
    if (OP == OP_VIADD_16X2) return __vadd2(x, a);
    if (OP == OP_VIMAX3_S32) return (unsigned)__vimax3_s32((int)x, (int)a, (int)b);
    if (OP == OP_VIADDMAX_S32) return (unsigned)__viaddmax_s32((int)x, (int)a, (int)b);
    if (OP == OP_VIMIN3_U32) return __vimin3_u32(x, a, b);
    if (OP == OP_IADD3) return x + a + b;
    if (OP == OP_LOP3) return (x & a) ^ b;
    if (OP == OP_PRMT) return __byte_perm(x, a, b & 0x7777);
    if (OP == OP_IMAD) return x * a + b;
    if (OP == OP_HSET2) {
        unsigned r;
        asm("set.eq.u32.f16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(a));
        return r ^ b;
    }
    if (OP == OP_SHFL) return __shfl_up_sync(0xffffffffu, x, 1) + a;
    if (OP == OP_LDS) return sm[(x + a) & 1023];
    if (OP == OP_VIMNMX_PRED) {
        bool p, q;
        unsigned r = __vibmax_s16x2(x, a, &p, &q);
        return r + (p ? 1u : 0u) + (q ? b : 0u);
    }
    return x;
}

template <int OP>
__global__ void __launch_bounds__(256) bench_kernel(unsigned *out, long long *cycles, unsigned a, unsigned b)
{
    __shared__ unsigned sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 7u;
    __syncthreads();
    unsigned x[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) x[c] = threadIdx.x * 31u + c + blockIdx.x;
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            if (OP == OP_MIX_DPX_IMAD) {
                if (c & 1)
                    x[c] = x[c] * a + b;
                else
                    x[c] = __vimax3_s16x2(x[c], a, b + it);
            } else if (OP == OP_MIX_DPX_LOP) {
                if (c & 1)
                    x[c] = (x[c] & a) ^ (b + it);
                else
                    x[c] = __vimax3_s16x2(x[c], a, b + it);
            } else {
                x[c] = step<OP>(x[c], a + c, b + it, sm, lane);
            }
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int c = 0; c < NCH; c++) acc ^= x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, unsigned *d_out, long long *d_cyc, int clock_khz)
{
    const int blocks = sms * 4, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    bench_kernel<OP><<<blocks, threads>>>(d_out, d_cyc, 0x00030005u, 0x00010002u);  // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    double cyc_avg = 0;
    for (int rep = 0; rep < 12; rep++) {
        cudaEventRecord(e0);
        bench_kernel<OP><<<blocks, threads>>>(d_out, d_cyc, 0x00030005u, 0x00010002u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) {
            best = ms;
            std::vector<long long> h(blocks);
            cudaMemcpy(h.data(), d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
            double s = 0;
            for (auto v : h) s += (double)v;
            cyc_avg = s / blocks;
        }
    }
    const double ops = (double)blocks * threads * (double)ITERS * NCH;
    const double per_sm_clk = (double)(4 * threads) * ITERS * NCH / cyc_avg;  // 4 resident blocks per SM
    printf("{\"op\": \"%s\", \"gops\": %.1f, \"ms\": %.4f, \"results_per_clk_per_sm\": %.2f, \"avg_block_cycles\": %.0f, "
           "\"implied_mhz\": %.0f}\n",
           op_names[OP], ops / best * 1e-6, best, per_sm_clk, cyc_avg, cyc_avg / best * 1e-3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    (void)clock_khz;
}

// ---- HBM write-only / read-only / copy bandwidth (the fill kernels are pure writers)
__global__ void bw_write_kernel(uint4 *dst, size_t n16, unsigned v)
{
    const uint4 val = make_uint4(v, v + 1, v + 2, v + 3);
    for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < n16; x += (size_t)gridDim.x * blockDim.x) dst[x] = val;
}
__global__ void bw_read_kernel(const uint4 *src, size_t n16, unsigned *out)
{
    unsigned acc = 0;
    for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < n16; x += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = src[x];
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}
__global__ void bw_copy_kernel(uint4 *dst, const uint4 *src, size_t n16)
{
    for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < n16; x += (size_t)gridDim.x * blockDim.x) dst[x] = src[x];
}
static void run_bandwidth(int sms, unsigned *d_out)
{
    const size_t bytes = (size_t)8 << 30, n16 = bytes / 16;
    uint4 *a = nullptr, *b = nullptr;
    if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, bytes) != cudaSuccess) {
        fprintf(stderr, "bandwidth test: allocation failed\n");
        return;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * 16, threads = 512;
    const char *names[3] = {"hbm_write_only", "hbm_read_only", "hbm_copy"};
    for (int k = 0; k < 3; k++) {
        float best = 1e30f;
        for (int rep = 0; rep < 6; rep++) {
            cudaEventRecord(e0);
            if (k == 0) bw_write_kernel<<<blocks, threads>>>(a, n16, (unsigned)rep);
            if (k == 1) bw_read_kernel<<<blocks, threads>>>(a, n16, d_out);
            if (k == 2) bw_copy_kernel<<<blocks, threads>>>(b, a, n16);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        const double moved = (k == 2) ? 2.0 * bytes : (double)bytes;
        printf("{\"op\": \"%s\", \"GBps\": %.1f, \"ms\": %.3f, \"bytes_moved\": %.0f}\n", names[k], moved / best * 1e-6, best, moved);
    }
    cudaFree(a);
    cudaFree(b);
}

int main()
{
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) {
        fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    unsigned *d_out;
    long long *d_cyc;
    cudaMalloc(&d_out, sizeof(unsigned) * sms * 4 * 256);
    cudaMalloc(&d_cyc, sizeof(long long) * sms * 4);
    run<OP_VIMAX3_S16X2>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_VIADDMAX_S16X2_RELU>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_VIADD_16X2>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_VIMNMX_PRED>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_VIMAX3_S32>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_VIADDMAX_S32>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_VIMIN3_U32>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_IADD3>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_LOP3>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_PRMT>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_IMAD>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_HSET2>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_MIX_DPX_IMAD>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_MIX_DPX_LOP>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_SHFL>(sms, d_out, d_cyc, prop.clockRate);
    run<OP_LDS>(sms, d_out, d_cyc, prop.clockRate);
    run_bandwidth(sms, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
