// Host-side helpers shared by the C-ABI entry points: error reporting, argument checks and
// TMA tensor-map encoding (cuTensorMapEncodeTiled is fetched through the runtime so the
// library has no link-time dependency on libcuda and loads on a box without a GPU).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "b200tta.h"

namespace b200 {

void set_last_error(const char* fmt, ...);

#define B200_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            b200::set_last_error(__VA_ARGS__);  \
            return B200TTA_EINVAL;              \
        }                                       \
    } while (0)

// use right after a <<<>>> launch: counts it and surfaces launch errors
#define B200_LAUNCHED()                   \
    do {                                  \
        b200::count_launch(1);            \
        B200_CUDA(cudaGetLastError());    \
    } while (0)

#define B200_CUDA(call)                                                                        \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            b200::set_last_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return B200TTA_ECUDA;                                                              \
        }                                                                                      \
    } while (0)

// process-wide count of kernels this library has launched (b200tta_launch_count)
void count_launch(int n = 1);

// returns 0 when the current device is sm_100 (B200); EARCH otherwise.  Cached per process.
int require_sm100();
int sm_count();

// 2-D bf16 tensor map, 128-byte swizzle.  dims are {inner, outer} in elements; the box is
// {box_inner (<= 64), box_outer (<= 256)}.  Out-of-bounds elements read as zero.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer);
// 3-D variant: dims {d0 (inner), d1, d2}, strides in bytes for d1, d2.
int make_tmap_2d_f32_plain(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                           uint32_t box_inner, uint32_t box_outer);
int make_tmap_3d_f32_plain(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                           uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2);
int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace b200
