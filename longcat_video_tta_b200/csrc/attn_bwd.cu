// placeholder until the tcgen05 backward lands (fails loudly; there is no fallback)
#include "host_common.h"
extern "C" int b200tta_attn_bwd(void*, int64_t, void*, int64_t, void*, int64_t, const void*, int64_t, const void*,
                                int64_t, const float*, float*, const void*, int64_t, const void*, int64_t, const void*,
                                int64_t, int32_t, int32_t, int32_t, float, const b200tta_attn_seg*, int32_t,
                                b200tta_stream_t) {
    b200::set_last_error("attn_bwd: not built yet");
    return B200TTA_EINVAL;
}
