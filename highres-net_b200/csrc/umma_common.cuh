// Shared by the tcgen05 kernels (conv3x3_umma.cu, decoder_umma.cu): descriptor helpers and the host-side
// TMA tensor-map encoder for bf16 NHWC activation tensors.
#pragma once
#include "internal.h"
#include "ptx.cuh"

#include <mutex>
#include <utility>

namespace hrn {

// K-major SWIZZLE_128B shared-memory descriptor: only the low word (start address) varies.
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc(uint32_t lo) { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return (saddr >> 4) | (1u << 16); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Guards the lazily initialised per-device statics below (two handles on two host threads are legal).
inline std::mutex& lazy_init_mutex() {
    static std::mutex mu;
    return mu;
}

inline EncodeTiledFn get_encode_fn() {
    std::lock_guard<std::mutex> lock(lazy_init_mutex());
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 4-D map over a bf16 NHWC tensor (dims fastest first: C, W, H, N); box = 64 channels x box_pixels x 1 x 1,
// SWIZZLE_128B, out-of-bounds elements read as zero (that is the conv padding).
inline int encode_nhwc_map(CUtensorMap* map, const void* base, int channels, int W, int H, int images, int box_pixels,
                           int image_stride = 1, int box_images = 1) {
    EncodeTiledFn encode = get_encode_fn();
    if (encode == nullptr) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return -1;
    }
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(channels), static_cast<cuuint64_t>(W),
                                static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(images)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(channels) * 2, static_cast<cuuint64_t>(W) * channels * 2,
                                   static_cast<cuuint64_t>(H) * W * channels * 2 * image_stride};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(box_pixels), 1, static_cast<cuuint32_t>(box_images)};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (dims %d x %d x %d x %d)", (int)r, channels, W, H, images);
        return -1;
    }
    return 0;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: one flag per device and call site, so a process
// that drives several GPUs (one handle per device) opts every one of them in.
template <typename K>
inline int allow_dynamic_smem(K kernel, int bytes, bool (&done)[64]) {
    int dev = 0;
    HRN_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) {
        set_error("device index %d out of range", dev);
        return -1;
    }
    std::lock_guard<std::mutex> lock(lazy_init_mutex());
    if (!done[dev]) {
        HRN_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        done[dev] = true;
    }
    return 0;
}

// Launch with programmatic dependent launch enabled: the kernel's prologue (barrier init, TMEM allocation, weight
// loads) overlaps the tail of the previous kernel in the stream; the kernel itself calls ptx::pdl_wait() before it
// reads or writes any tensor the previous kernel may still be using.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, int cluster, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cluster > 1) {                      // thread-block clusters of `cluster` consecutive CTAs
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = cluster;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = 1;
        cfg.numAttrs = 2;
    }
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace hrn
