// Cost of tcgen05.commit / fences / mbarrier polls interleaved with an N=192 MMA stream.
#include "ptx.cuh"
#include <cstdio>
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t mk(uint32_t lo) { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; }
// mode bit0: commit to a scratch barrier every 12 MMAs; bit1: a second commit; bit2: tcgen05.fence::after every 12;
// bit3: an (already complete) mbarrier try_wait every 12; bit4: commits target barriers that a second warp consumes
__global__ void __launch_bounds__(128, 1) probe(long long* out, int mode) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_s = base, b_s = base + 32768, bar = base + 98304, scratch = bar + 8, scratch2 = bar + 16, done = bar + 24, slot = bar + 32;
    uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    for (int i = threadIdx.x; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) {
        ptx::mbar_init(bar, 1); ptx::mbar_init(scratch, 1); ptx::mbar_init(scratch2, 1); ptx::mbar_init(done, 1);
        ptx::fence_barrier_init();
        ptx::mbar_arrive(done);     // phase 0 of `done` is complete: try_wait(done, 0) succeeds immediately
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) ptx::tmem_alloc<512>(slot);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
    if (threadIdx.x < 32) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, 192);
        const uint32_t a_lo = (a_s >> 4) | (1u << 16), b_lo = (b_s >> 4) | (1u << 16);
        const long long t0 = clock64();
        if (ptx::elect_one()) {
            for (int rep = 0; rep < 400; ++rep) {
                if (mode & 8) ptx::mbar_wait(done, 0, 1);
                if (mode & 4) ptx::tc_fence_after();
#pragma unroll
                for (int k = 0; k < 12; ++k)
                    ptx::umma_bf16(tmem + (rep & 1) * 256, mk(a_lo + (k / 4) * 8 + (k % 4) * 2), mk(b_lo + k * 2), idesc, 1u);
                if (mode & 1) ptx::umma_commit(scratch);
                if (mode & 2) ptx::umma_commit(scratch2);
            }
            ptx::umma_commit(bar);
        }
        __syncwarp();
        ptx::mbar_wait(bar, 0, 9);
        const long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}
int main() {
    const int ctas = 148, smem = 98304 + 1024 + 128;
    long long* d; cudaMalloc(&d, sizeof(long long) * ctas);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode : {0, 1, 3, 4, 8, 12, 15}) {
        probe<<<ctas, 128, smem>>>(d, mode);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
        printf("mode %2d (commit=%d commit2=%d fence=%d poll=%d): %.0f cycles per 12 MMAs (floor 1152) [%s]\n", mode, mode & 1,
               (mode >> 1) & 1, (mode >> 2) & 1, (mode >> 3) & 1, avg / 400.0, cudaGetErrorString(e));
    }
    return 0;
}
