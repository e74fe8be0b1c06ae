// tcgen05.mma time vs accumulator column offset in TMEM (M = 128, SS, N = 64/128/192).
#include "ptx.cuh"
#include <cstdio>
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t mk(uint32_t lo) { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; }
template <int N>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int col, int col2) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_s = base, b_s = base + 32768, bar = base + 98304, slot = bar + 16;
    uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    for (int i = threadIdx.x; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) ptx::tmem_alloc<512>(slot);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
    if (threadIdx.x < 32) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, N);
        const uint32_t a_lo = (a_s >> 4) | (1u << 16), b_lo = (b_s >> 4) | (1u << 16);
        long long t0 = clock64();
        if (ptx::elect_one()) {
            for (int rep = 0; rep < 100; ++rep) {
#pragma unroll
                for (int k = 0; k < 24; ++k)
                    ptx::umma_bf16(tmem + ((k / 12) & 1 ? col2 : col), mk(a_lo + (k % 4) * 2), mk(b_lo + (k % 12) * 2), idesc, 1u);
            }
            ptx::umma_commit(bar);
        }
        __syncwarp();
        ptx::mbar_wait(bar, 0, 9);
        long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}
template <int N> void run(int col, int col2) {
    const int ctas = 148, smem = 98304 + 1024 + 64;
    long long* d; cudaMalloc(&d, sizeof(long long) * ctas);
    cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<N><<<ctas, 128, smem>>>(d, col, col2);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
    printf("N=%3d cols %3d/%3d (12 MMAs each, alternating): %.1f cycles/MMA [%s]\n", N, col, col2, avg / 2400.0, cudaGetErrorString(e));
    cudaFree(d);
}
int main() {
    for (int c : {0, 64, 128, 192, 256, 320}) run<192>(c, c);
    for (int c : {0, 64, 128}) run<192>(c, c + 64);     // what the conv does: next row's run starts one slot later
    run<192>(0, 256); run<192>(0, 128); run<192>(0, 192);
    for (int c : {0, 64, 192, 448}) run<64>(c, c);
    run<128>(0, 0); run<128>(64, 64); run<128>(384, 384);
    return 0;
}
