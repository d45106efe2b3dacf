// shipsim_host.h -- host-side helpers of the C ABI (no CUDA types)
#pragma once
#include <cstddef>
#include <cstdint>

namespace shipsim {

// obs[row] = [frames[row] | frames[row + N]] for row in [row_begin, row_end), 16 floats per frame; where cut[row] != 0
// (cut may be NULL) the first half is 16 x -1 instead: the reset observation of ShipEnv.reset (ship_env.py:180-184).
// `frames` holds the frame of the state before the call for every env (N frames) followed by one frame per env-step.
void assemble_history_rows(float *obs, const float *frames, const uint8_t *cut, size_t row_begin, size_t row_end, size_t N);

}  // namespace shipsim
