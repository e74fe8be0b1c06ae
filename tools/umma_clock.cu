// Sustained tensor-pipe rate and SM clock under load: N=192 MMAs back to back on all SMs for a long time;
// clock64 vs globaltimer gives the effective SM frequency while the tensor pipe is saturated.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I highres-net_b200/csrc tools/umma_clock.cu -o tools/bin/umma_clock
#include "ptx.cuh"
#include <cstdio>
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t mk(uint32_t lo) { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; }
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void __launch_bounds__(128, 1) burn(long long* out, int reps) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_s = base, b_s = base + 32768, bar = base + 98304, slot = bar + 16;
    uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    for (int i = threadIdx.x; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_barrier_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) ptx::tmem_alloc<512>(slot);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
    if (threadIdx.x < 32) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, 192);
        const uint32_t a_lo = (a_s >> 4) | (1u << 16), b_lo = (b_s >> 4) | (1u << 16);
        const long long c0 = clock64();
        const unsigned long long g0 = gtime();
        uint32_t phase = 0;
        for (int rep = 0; rep < reps; ++rep) {
            if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < 48; ++k)
                    ptx::umma_bf16(tmem + (k & 1) * 256, mk(a_lo + ((k / 4) % 3) * 8 + (k % 4) * 2),
                                   mk(b_lo + (k % 12) * 2), idesc, 1u);
                ptx::umma_commit(bar);
            }
            __syncwarp();
            ptx::mbar_wait(bar, phase, 9);
            phase ^= 1;
        }
        const long long c1 = clock64();
        const unsigned long long g1 = gtime();
        if (threadIdx.x == 0) {
            out[2 * blockIdx.x] = c1 - c0;
            out[2 * blockIdx.x + 1] = (long long)(g1 - g0);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}
int main() {
    const int ctas = 148, smem = 98304 + 1024 + 64, reps = 20000;   // ~ 20000 * 48 * 101 cycles ~ 97 M cycles ~ 50-75 ms
    long long* d;
    cudaMalloc(&d, sizeof(long long) * 2 * ctas);
    cudaFuncSetAttribute(burn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int launch = 0; launch < 12; ++launch) {
        burn<<<ctas, 128, smem>>>(d, reps);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[296];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double cyc = 0, ns = 0;
        for (int i = 0; i < ctas; ++i) {
            cyc += h[2 * i];
            ns += h[2 * i + 1];
        }
        cyc /= ctas;
        ns /= ctas;
        const double flops = 2.0 * 128 * 192 * 16 * 48.0 * reps * ctas;
        printf("launch %2d: %.1f cycles/MMA, %.1f ms, SM clock under load %.0f MHz, %.0f TFLOP/s [%s]\n", launch,
               cyc / (48.0 * reps), ns / 1e6, cyc / ns * 1e3, flops / ns / 1e3, cudaGetErrorString(e));
    }
    return 0;
}
