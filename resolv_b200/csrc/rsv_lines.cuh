// rsv_lines.cuh — batched LineTest (ray versus cell grid). Filled in after the frame path is parity-green.
#pragma once
#include "rsv_device.cuh"

namespace rsv {
struct LineState {};
inline void lines_free(LineState&) {}
}  // namespace rsv
