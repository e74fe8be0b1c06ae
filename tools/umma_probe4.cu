// Row-pattern replay: each row = 3 x N=64 first-step MMAs + 11 k-steps on accumulator slots s_lo..s_lo+2 (mod 8,
// split in two MMAs at the wrap).  Patterns of s_lo sequences show which transitions stall the tensor pipe.
#include "ptx.cuh"
#include <cstdio>
#include <vector>
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t mk(uint32_t lo) { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; }
__constant__ int c_pat[16];
__global__ void __launch_bounds__(128, 1) replay(long long* out, int rows, int plen, int first_step, int b_adv, int a_rot) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_s = base, a_s = base + 73728, bar = a_s + 17408, slot = bar + 16;
    uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    for (int i = threadIdx.x; i < (73728 + 17408) / 4; i += 128) reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) ptx::tmem_alloc<512>(slot);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
    if (threadIdx.x < 32) {
        long long t0 = clock64();
        if (ptx::elect_one()) {
            const uint32_t ib = ptx::umma_idesc_bf16(128, 0);
            const uint32_t a_lo0 = (a_s >> 4) | (1u << 16), b_lo0 = (b_s >> 4) | (1u << 16);
            for (int q = 0; q < rows; ++q) {
                const uint32_t s_lo = c_pat[q % plen];
                const int w0 = min(3, 8 - (int)s_lo), w1 = 3 - w0;
                const uint32_t d0 = tmem + s_lo * 64, d1 = tmem;
                const uint32_t id0 = ib | ((uint32_t)(w0 * 64 >> 3) << 17), id1 = ib | ((uint32_t)(w1 * 64 >> 3) << 17);
                uint64_t ad = mk(a_lo0 + (a_rot ? (q % 8) * (17408 / 16) : 0)), bd0 = mk(b_lo0), bd1 = bd0 + w0 * 512;
                if (first_step == 1) {
                    for (int b = 0; b < 3; ++b)
                        ptx::umma_bf16(tmem + ((s_lo + b) % 8) * 64, ad, bd0 + b * 512, ib | (8u << 17), b == 2 ? 0u : 1u);
                } else {
                    ptx::umma_bf16(d0, ad, bd0, id0, 1u);
                    if (w1 > 0) ptx::umma_bf16(d1, ad, bd1, id1, 1u);
                }
#pragma unroll
                for (int step = 1; step < 12; ++step) {
                    ad += 2; if (b_adv) { bd0 += (step & 3) ? 2u : (24576 / 16 - 6); bd1 += (step & 3) ? 2u : (24576 / 16 - 6); }
                    ptx::umma_bf16(d0, ad, bd0, id0, 1u);
                    if (w1 > 0) ptx::umma_bf16(d1, ad, bd1, id1, 1u);
                }
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
void run(std::vector<int> pat, int first_step, int b_adv, int a_rot, int big_smem) {
    const int ctas = 148, rows = 480, smem = 73728 + (big_smem ? 8 : 1) * 17408 + 1024 + 64;
    long long* d; cudaMalloc(&d, sizeof(long long) * ctas);
    cudaMemcpyToSymbol(c_pat, pat.data(), pat.size() * sizeof(int));
    cudaFuncSetAttribute(replay, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    replay<<<ctas, 128, smem>>>(d, rows, (int)pat.size(), first_step, b_adv, a_rot);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
    double expect = 0;
    for (int s : pat) expect += (first_step ? 144 : 0) + (first_step ? 11 : 12) * (s <= 5 ? 96.0 : 112.0);
    printf("first_step=%d b_adv=%d a_rot=%d big_smem=%d pattern", first_step, b_adv, a_rot, big_smem); for (int s : pat) printf(" %d", s);
    printf(" : %.0f cycles/row (expected %.0f) [%s]\n", avg / rows, expect / pat.size(), cudaGetErrorString(e));
    cudaFree(d);
}
int main() {
    run({0, 1, 2, 3, 4, 5}, 0, 1, 0, 0); run({0, 1, 2, 3, 4, 5}, 0, 0, 0, 0); run({0, 1, 2, 3, 4, 5}, 0, 1, 0, 1);
    run({0, 1, 2, 3, 4, 5}, 0, 1, 1, 1); run({0, 1, 2, 3, 4, 5}, 1, 1, 1, 1); run({0, 1, 2, 3, 4, 5, 6, 7}, 1, 1, 1, 1);
    run({0, 1, 2, 3, 4, 5, 6, 7}, 1, 0, 1, 1); run({0}, 0, 1, 0, 1); run({0}, 0, 1, 1, 1);
    return 0;
}
