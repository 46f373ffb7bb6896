// Error reporting, device checks and TMA tensor-map encoding for libb200tta.so.
#include "host_common.h"

#include <atomic>
#include <mutex>
#include <string.h>

namespace b200 {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int g_arch_state = 0;  // 0 unknown, 1 ok, -1 not sm_100
static int g_sm_count = 0;

int require_sm100() {
    if (g_arch_state == 0) {
        int dev = 0, major = 0, minor = 0, sms = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
            cudaGetLastError();
            set_last_error("b200tta: no usable CUDA device (this library has no CPU fallback)");
            return B200TTA_EARCH;
        }
        g_sm_count = sms;
        g_arch_state = (major == 10 && minor == 0) ? 1 : -1;
        if (g_arch_state < 0) set_last_error("b200tta: device is sm_%d%d, need sm_100 (B200); no fallback path", major, minor);
    }
    if (g_arch_state < 0) {
        set_last_error("b200tta: device is not sm_100 (B200); no fallback path");
        return B200TTA_EARCH;
    }
    return B200TTA_OK;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() { return g_sm_count > 0 ? g_sm_count : 148; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

static int encode(CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_last_error("b200tta: cuTensorMapEncodeTiled not available from the driver");
        return B200TTA_ECUDA;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("b200tta: cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu] stride1 %llu box [%u,%u,%u] base %p",
                       (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                       (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)strides[0], box[0], box[1],
                       rank > 2 ? box[2] : 0, base);
        return B200TTA_EINVAL;
    }
    return B200TTA_OK;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    return encode(out, base, 2, dims, strides, box);
}

// fp32, no swizzle: the destination of TMA reduce-adds (cp.reduce.async.bulk.tensor), box rows contiguous in shared memory
int make_tmap_2d_f32_plain(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                           uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_last_error("b200tta: cuTensorMapEncodeTiled not available from the driver");
        return B200TTA_ECUDA;
    }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("b200tta: cuTensorMapEncodeTiled (f32) failed (%d): dims [%llu,%llu] stride %llu box [%u,%u] base %p", (int)r,
                       (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner,
                       box_outer, base);
        return B200TTA_EINVAL;
    }
    return B200TTA_OK;
}

int make_tmap_3d_f32_plain(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                           uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_last_error("b200tta: cuTensorMapEncodeTiled not available from the driver");
        return B200TTA_ECUDA;
    }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {box0, box1, box2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("b200tta: cuTensorMapEncodeTiled (f32, 3-D) failed (%d): dims [%llu,%llu,%llu] box [%u,%u,%u] base %p", (int)r,
                       (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, box0, box1, box2, base);
        return B200TTA_EINVAL;
    }
    return B200TTA_OK;
}

int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2) {
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {box0, box1, box2};
    return encode(out, base, 3, dims, strides, box);
}

}  // namespace b200

namespace b200 { unsigned long long launches(); }
extern "C" int b200tta_version(void) { return 100; }
extern "C" int64_t b200tta_launch_count(void) { return (int64_t)b200::launches(); }
extern "C" const char* b200tta_last_error(void) { return b200::g_err; }
