// Micro-benchmark: issue rate of tcgen05.mma (cta_group::1, kind::f16, M = 128, SS operands, SWIZZLE_128B)
// as a function of N and of the A-operand start-address shift used by the conv kernel (kx * 128 B).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I highres-net_b200/csrc tools/umma_probe.cu -o tools/bin/umma_probe
#include "ptx.cuh"
#include <cstdio>
#include <cstdlib>

constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t mk(uint32_t lo) { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; }

// mode bits: shift pattern of A per MMA: 0 = always aligned, 1 = cycles 0,128,256 B like the conv taps
template <int N, int SHIFTS, int ALT>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int reps, int spin) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_s = base, b_s = base + 32768, bar = base + 32768 + 65536, slot = bar + 16;
    uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
    for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += 128) reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) ptx::tmem_alloc<512>(slot);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (slot - base));
    if (threadIdx.x < 32) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, N);
        const uint32_t a_lo = (a_s >> 4) | (1u << 16), b_lo = (b_s >> 4) | (1u << 16);
        long long t0 = 0, t1 = 0;
        uint32_t phase = 0;
        for (int rep = 0; rep < reps + 1; ++rep) {
            if (rep == 1) t0 = clock64();
            if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < 48; ++k) {
                    const int kx = SHIFTS ? (k / 4) % 3 : 0, j = k % 4;
                    ptx::umma_bf16(tmem + (ALT ? (k & 1) * 256 : 0), mk(a_lo + kx * 8 + j * 2), mk(b_lo + (k % 12) * 2), idesc, 1u);
                    if (spin > 0 && (k % 12) == 11) { const long long c0 = clock64(); while (clock64() - c0 < spin) {} }
                }
                ptx::umma_commit(bar);
            }
            __syncwarp();
            ptx::mbar_wait(bar, phase, 9);
            phase ^= 1;
        }
        t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0);
    }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

template <int N, int SHIFTS, int ALT>
void run(int ctas, int spin = 0) {
    long long* d; cudaMalloc(&d, sizeof(long long) * ctas);
    const int smem = 32768 + 65536 + 1024 + 64, reps = 200;
    cudaFuncSetAttribute(probe<N, SHIFTS, ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<N, SHIFTS, ALT><<<ctas, 128, smem>>>(d, reps, spin);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
    printf("spin=%4d N=%3d shifted_A=%d alt_acc=%d ctas=%3d : %.1f cycles/MMA (floor %d)  smem bytes/MMA %d -> %.1f B/cycle  [%s]\n", spin, N, SHIFTS, ALT, ctas,
           avg / (reps * 48.0), N / 2, 4096 + N * 32, (4096 + N * 32) / (avg / (reps * 48.0)), cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int spin : {0, 100, 200, 300, 400, 600, 800, 1200}) run<192, 1, 0>(148, spin);
    return 0;
}
