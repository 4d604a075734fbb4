// tcgen05 rollout kernel -- placeholder until the UMMA path lands (next commit).
#include "common.cuh"
#include "pe_kernels.h"
namespace mb {
size_t tc_pack_bytes(const mb_mlp_desc&) { return 0; }
int launch_tc_pack(const PeDev&, void*, cudaStream_t) { set_error("tcgen05 path not built"); return MB_EUNSUPPORTED; }
int launch_pe_rollout_tc(const PeDev&, const void*, const float*, const float*, const uint8_t*, const float*,
                         int64_t, int, float*, float*, float*, void*, cudaStream_t) {
    set_error("tcgen05 path not built");
    return MB_EUNSUPPORTED;
}
}  // namespace mb
