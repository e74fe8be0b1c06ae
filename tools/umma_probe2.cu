// Replays the conv3x3 kernel's MMA stream for interior rows (no TMA, no epilogue) to find what the tensor pipe
// itself sustains: per row 3 x N=64 first-step MMAs + 11 x N=192 (split N=128+64 / 64+128 at the TMEM wrap).
#include "ptx.cuh"
#include <cstdio>
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t mk(uint32_t lo) { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; }

// MODE 0: fixed accumulator column 0, same B for all steps      MODE 1: rotating accumulator slots (with wrap split)
// MODE 2: MODE 1 + distinct B tile per step (72 KB footprint)   MODE 3: MODE 2 + rotating A buffers (8 x 17 KB)
template <int MODE>
__global__ void __launch_bounds__(128, 1) replay(long long* out, int rows, long long* trace) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_s = base, a_s = base + 73728, bar = a_s + 8 * 17408, slot = bar + 16;
    uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    for (int i = threadIdx.x; i < (73728 + 8 * 17408) / 4; i += 128) reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u;
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
            for (int q = 2; q < rows + 2; ++q) {
                if (blockIdx.x == 0 && q >= 202 && q < 234) trace[q - 202] = clock64();
                const uint32_t t_lo = q - 2, s_lo = MODE == 4 ? t_lo % 6 : (MODE == 5 || MODE == 6) ? 6 + (t_lo & 1) : (MODE >= 1 ? t_lo % 8 : 0);
                const int w0 = MODE >= 1 ? min(3, 8 - (int)s_lo) : 3, w1 = 3 - w0;
                const uint32_t d0 = tmem + s_lo * 64, d1 = tmem;
                const uint32_t id0 = ib | ((uint32_t)(w0 * 64 >> 3) << 17), id1 = ib | ((uint32_t)(w1 * 64 >> 3) << 17);
                uint64_t ad = mk(a_lo0 + (MODE >= 3 ? (q % 8) * (17408 / 16) : 0));
                uint64_t bd0 = mk(b_lo0), bd1 = bd0 + w0 * 512;
                for (int b = 0; b < 3; ++b)
                    ptx::umma_bf16(tmem + (MODE >= 1 ? ((t_lo + b) % 8) * 64 : b * 64), ad, bd0 + b * 512, ib | (8u << 17), b == 2 ? 0u : 1u);
                if (MODE == 6) {
                    uint64_t ad2 = ad;
#pragma unroll
                    for (int step = 1; step < 12; ++step) { ad += 2; ptx::umma_bf16(d0, ad, bd0, id0, 1u); }
#pragma unroll
                    for (int step = 1; step < 12; ++step) { ad2 += 2; ptx::umma_bf16(d1, ad2, bd1, id1, 1u); }
                } else
#pragma unroll
                for (int step = 1; step < 12; ++step) {
                    ad += 2;
                    if (MODE >= 2) { bd0 += (step & 3) ? 2u : (24576 / 16 - 6); bd1 += (step & 3) ? 2u : (24576 / 16 - 6); }
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

template <int MODE> void run() {
    const int ctas = 148, rows = 400, smem = 73728 + 8 * 17408 + 1024 + 64;
    long long* d; cudaMalloc(&d, sizeof(long long) * ctas);
    cudaFuncSetAttribute(replay<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* tr; cudaMalloc(&tr, 32 * 8);
    replay<MODE><<<ctas, 128, smem>>>(d, rows, tr);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
    printf("mode %d: %.0f cycles per row (MMA floor 12 x 96 = 1152) [%s]\n", MODE, avg / rows, cudaGetErrorString(e));
    long long ht[32]; cudaMemcpy(ht, tr, sizeof(ht), cudaMemcpyDeviceToHost);
    printf("   issue-time deltas per row (s_lo = (q-2)%%8 starting at q=202 -> s_lo 0):");
    for (int i = 1; i < 25; ++i) printf(" %lld", ht[i] - ht[i - 1]);
    printf("\n");
    cudaFree(d); cudaFree(tr);
}
int main() { run<1>(); run<4>(); return 0; }
