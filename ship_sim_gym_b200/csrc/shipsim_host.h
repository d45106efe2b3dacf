// shipsim_host.h -- host-side helpers of the C ABI (no CUDA types)
#pragma once
#include <cstddef>
#include <cstdint>

namespace shipsim {

// obs[row] = [frames[row] | frames[row + N]] for row in [row_begin, row_end), 16 floats per frame; where cut[row] != 0
// (cut may be NULL) the first half is 16 x -1 instead: the reset observation of ShipEnv.reset (ship_env.py:180-184).
// `frames` holds the frame of the state before the call for every env (N frames) followed by one frame per env-step.
void assemble_history_rows(float *obs, const float *frames, const uint8_t *cut, size_t row_begin, size_t row_end, size_t N);

// The host half of the compacted wire format (compact_frames_kernel): rebuild observation rows, rewards and done flags
// of `kc` steps for the env blocks [blk_begin, blk_end) (32 envs each) from 16-byte records + the stream of changed
// values.  `cur` = one 16-float frame per env: the frame before the first of these steps on entry, the frame after the
// last one on exit.  rec / off are indexed from the chunk's first step; obs / rew / done (any may be NULL) from the
// chunk's first row.  history = 1: rows are frames; 2: [previous frame | frame], with 16 x -1 as the previous frame of a
// row whose done flag is set when cut_on_done (auto-reset: ship_env.py:180-184).
void expand_delta_rows(float *obs, float *rew, uint8_t *done, const uint32_t *rec, const uint32_t *off, const float *var, float *cur,
                       int kc, size_t N, size_t blk_begin, size_t blk_end, float step_penalty, bool cut_on_done, int history);

}  // namespace shipsim
