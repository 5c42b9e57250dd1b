// kernel_windowed.cuh -- persistent screen-tile kernel marching through shared-memory windows
// staged by TMA (placeholder until the windowed path lands; AUTO resolves to the direct kernel).
#pragma once

#include "march_device.cuh"

namespace vr {

struct WindowedState { int unused = 0; };

inline bool windowed_supported(const FrameConsts&, int) { return false; }
inline const char* windowed_last_error() { return "windowed kernel not built"; }
inline void windowed_invalidate(WindowedState&) {}
inline void windowed_release(WindowedState&) {}
inline int launch_windowed(WindowedState&, const FrameConsts&, const void*, int, uint32_t, uint64_t,
                           float*, int, int, cudaStream_t, uint32_t*) { return -1; }

}  // namespace vr
