// tools/microbench.cu — B200 hardware probes that size the matcher's epilogue budget:
//   (1) tcgen05.mma.kind::i8 issue rate with no epilogue (dense int8 peak as this chip delivers it)
//   (2) tcgen05.ld (TMEM -> registers) throughput
//   (3) CUDA-core op rates for the candidate epilogue instructions (IMAD, VIADDMNMX, VIMNMX3, REDUX, ...)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../metricsfm_b200/csrc/sm100_ptx.cuh"

using namespace msfm;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// ------------------------------------------------------------------ (1) MMA issue rate
template <int N, int STRIPS>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, unsigned long long *cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    // A: STRIPS x 128 rows x 128 B, B: N rows x 128 B; contents irrelevant (zeros)
    for (int i = threadIdx.x; i < (STRIPS * 128 + N) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    ptx::fence_proxy_async();
    if (threadIdx.x < 32) ptx::tmem_alloc<512>(&tmem_ptr);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = ptx::make_idesc_i8(128, N, 0, 0);
        const uint32_t a_addr = ptx::smem_u32(smem), b_addr = a_addr + STRIPS * 128 * 128;
        const long long t0 = clock64();
        uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            const uint32_t buf = it & 1;
#pragma unroll
            for (int s = 0; s < STRIPS; ++s)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    ptx::mma_i8_ss(tmem + (buf * STRIPS + s) * N % 512, ptx::make_smem_desc_sw128(a_addr + s * 16384 + k * 32),
                                   ptx::make_smem_desc_sw128(b_addr + k * 32), idesc, k > 0);
            if ((it & 7) == 7) {  // bound the number of MMAs in flight
                ptx::mma_commit(&bar);
                ptx::mbar_wait(&bar, phase);
                phase ^= 1;
            }
        }
        ptx::mma_commit(&bar);
        ptx::mbar_wait(&bar, phase);
        const long long t1 = clock64();
        cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------ (2) TMEM load rate
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) ldtm_rate_kernel(int iters, unsigned long long *cycles, uint32_t *sink) {
    __shared__ uint32_t tmem_ptr;
    if (threadIdx.x < 32) ptx::tmem_alloc<512>(&tmem_ptr);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_ptr;
    const int warp = threadIdx.x >> 5;
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t r[32];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            ptx::tmem_ld_32x32b_x32(base + ((it * 4 + c) * 32) % 512, r);
            ptx::tmem_ld_wait();
            acc ^= r[0] ^ r[31];
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------ (3) ALU op rates
enum Op { OP_MIX_MAX3_IMAD = 100, OP_MIX_MAX3_VIADDMAX, OP_MIX_MAX_VIADDMAX, OP_MIX_MAX3_FFMA, OP_VIADDMAX_ONLY, OP_MAX_ONLY, OP_IMAD = 0, OP_IADD3, OP_VIADDMIN, OP_VIMIN3, OP_MIN, OP_REDUX, OP_LOP3, OP_ISETP_SEL, OP_FFMA, OP_IMAD_MIN, OP_SHFL, OP_LDS };
template <int OP>
__global__ void __launch_bounds__(512, 1) alu_rate_kernel(int iters, unsigned long long *cycles, int *sink, int seed) {
    __shared__ int sm[1024];
    sm[threadIdx.x] = threadIdx.x * seed;
    sm[threadIdx.x + 512] = seed;
    __syncthreads();
    int x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * (i + 1) + seed;
    int c = seed * 3 + 1, d = seed + 7;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = (float)x[i];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == OP_IMAD) x[i] = x[i] * c + d;
            if (OP == OP_IADD3) x[i] = x[i] + c + d;
            if (OP == OP_VIADDMIN) x[i] = __viaddmin_s32(x[i], c, d + i);
            if (OP == OP_VIMIN3) x[i] = __vimin3_s32(x[i], c + i, d);
            if (OP == OP_MIN) x[i] = min(x[i], c + i) + 1;
            if (OP == OP_REDUX) x[i] = __reduce_min_sync(0xFFFFFFFFu, x[i]) + i;
            if (OP == OP_LOP3) x[i] = (x[i] ^ c) & (d | i);
            if (OP == OP_ISETP_SEL) x[i] = (x[i] < c + it) ? d : x[i] + 1;
            if (OP == OP_FFMA) f[i] = fmaf(f[i], 1.0001f, 0.5f);
            if (OP == OP_IMAD_MIN) x[i] = min(x[i], (x[(i + 1) & 7] & 0xffff) * c + d);
            if (OP == OP_SHFL) x[i] = __shfl_xor_sync(0xFFFFFFFFu, x[i], 1) + i;
            if (OP == OP_LDS) x[i] = sm[(x[i] + i) & 1023];
            // pipe co-issue probes: even chains run one op type, odd chains another
            if (OP == OP_MIX_MAX3_IMAD) { if (i & 1) x[i] = x[i] * c + d; else x[i] = __vimax3_s32(x[i], c + i, d); }
            if (OP == OP_MIX_MAX3_VIADDMAX) { if (i & 1) x[i] = __viaddmax_s32(x[i], c, d + i); else x[i] = __vimax3_s32(x[i], c + i, d); }
            if (OP == OP_MIX_MAX_VIADDMAX) { if (i & 1) x[i] = __viaddmax_s32(x[i], c, d + i); else x[i] = max(x[i], c + i) ^ d; }
            if (OP == OP_MIX_MAX3_FFMA) { if (i & 1) f[i] = fmaf(f[i], 1.0001f, 0.5f); else x[i] = __vimax3_s32(x[i], c + i, d); }
            if (OP == OP_VIADDMAX_ONLY) x[i] = __viaddmax_s32(x[i], c, d + i);
            if (OP == OP_MAX_ONLY) x[i] = max(x[i], c + i) ^ d;
        }
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + (int)f[i];
    if (s == 0x7fffffff) sink[0] = s;
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <typename F>
static double run(const char *name, F launch, int blocks, unsigned long long *d_cycles, double work_per_block, const char *unit) {
    std::vector<unsigned long long> h(blocks);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(h.data(), d_cycles, blocks * 8, cudaMemcpyDeviceToHost));
    double cyc = 0; for (auto v : h) cyc += (double)v; cyc /= blocks;
    printf("%-34s %10.3f ms  %12.0f cyc/block  %9.2f %s/cyc/SM   chip %.3e %s/s\n", name, ms, cyc, work_per_block / cyc, unit,
           work_per_block * blocks / (ms * 1e-3), unit);
    return ms;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, cc %d.%d\n", prop.name, sms, prop.major, prop.minor);
    unsigned long long *d_cycles; CK(cudaMalloc(&d_cycles, 8 * 1024));
    int *d_sink; CK(cudaMalloc(&d_sink, 64));

    {   // (1) MMA
        const int iters = 20000;
        auto k1 = mma_rate_kernel<256, 1>; auto k2 = mma_rate_kernel<128, 2>; auto k3 = mma_rate_kernel<128, 1>;
        auto k4 = mma_rate_kernel<64, 4>; auto k5 = mma_rate_kernel<64, 2>;
        const int smem = 4 * 128 * 128 + 256 * 128 + 2048;
        CK(cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(k5, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        run("mma i8 M128 N256 (1 strip)", [&] { k1<<<sms, 128, smem>>>(iters, d_cycles); }, sms, d_cycles, 2.0 * 128 * 256 * 128 * iters, "op");
        run("mma i8 M128 N128 x2 strips", [&] { k2<<<sms, 128, smem>>>(iters, d_cycles); }, sms, d_cycles, 2.0 * 2 * 128 * 128 * 128 * iters, "op");
        run("mma i8 M128 N128 (1 strip)", [&] { k3<<<sms, 128, smem>>>(iters, d_cycles); }, sms, d_cycles, 2.0 * 128 * 128 * 128 * iters, "op");
        run("mma i8 M128 N64 x4 strips", [&] { k4<<<sms, 128, smem>>>(iters, d_cycles); }, sms, d_cycles, 2.0 * 4 * 128 * 64 * 128 * iters, "op");
        run("mma i8 M128 N64 x2 strips", [&] { k5<<<sms, 128, smem>>>(iters, d_cycles); }, sms, d_cycles, 2.0 * 2 * 128 * 64 * 128 * iters, "op");
        run("mma i8 M128 N256, 1 SM only", [&] { k1<<<1, 128, smem>>>(iters, d_cycles); }, 1, d_cycles, 2.0 * 128 * 256 * 128 * iters, "op");
    }
    {   // (2) LDTM
        const int iters = 20000;
        run("tcgen05.ld 32x32b.x32, 4 warps", [&] { ldtm_rate_kernel<4><<<sms, 128>>>(iters, d_cycles, (uint32_t *)d_sink); }, sms, d_cycles, 4.0 * 4 * 4096 * iters, "B");
        run("tcgen05.ld 32x32b.x32, 8 warps", [&] { ldtm_rate_kernel<8><<<sms, 256>>>(iters, d_cycles, (uint32_t *)d_sink); }, sms, d_cycles, 8.0 * 4 * 4096 * iters, "B");
        run("tcgen05.ld 32x32b.x32, 16 warps", [&] { ldtm_rate_kernel<16><<<sms, 512>>>(iters, d_cycles, (uint32_t *)d_sink); }, sms, d_cycles, 16.0 * 4 * 4096 * iters, "B");
    }
    {   // (3) ALU: 512 threads x 8 chains per iteration
        const int iters = 20000;
        const double w = 512.0 * 8 * iters;
#define ALU(OPNAME) run("alu " #OPNAME, [&] { alu_rate_kernel<OPNAME><<<sms, 512>>>(iters, d_cycles, d_sink, 3); }, sms, d_cycles, w, "lane-op")
        ALU(OP_IMAD); ALU(OP_IADD3); ALU(OP_VIADDMIN); ALU(OP_VIMIN3); ALU(OP_MIN); ALU(OP_REDUX); ALU(OP_LOP3);
        ALU(OP_ISETP_SEL); ALU(OP_FFMA); ALU(OP_IMAD_MIN); ALU(OP_SHFL); ALU(OP_LDS);
        ALU(OP_MIX_MAX3_IMAD); ALU(OP_MIX_MAX3_VIADDMAX); ALU(OP_MIX_MAX_VIADDMAX); ALU(OP_MIX_MAX3_FFMA); ALU(OP_VIADDMAX_ONLY); ALU(OP_MAX_ONLY);
    }
    return 0;
}
