// v2 SpMM: tables staged in shared memory, rows tiled (DESIGN.md "v2").  Placeholder until the tiled kernel lands.
#pragma once
#include "common.cuh"

namespace scs {

inline bool tiled_supported(scs_ctx*, int /*kind*/, int /*KP*/) { return false; }
inline int launch_spmm_tiled(scs_ctx*, int, int, int, const double*, int64_t, double*, int64_t, const int32_t*) {
    set_error("tiled SpMM variant not built");
    return 1;
}

}  // namespace scs
