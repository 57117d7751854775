// Control-rate (250 Hz) helpers that feed the synthesis path.
#pragma once
#include "common.cuh"

namespace b200ddsp {

// NoteRelease (reference modules/sub_modules.py:1174-1188): an RNN over the frames of one voice
// whose cell, F0ProcessorCell.call (:1138-1171), holds the last played note for `release_frames`
// frames after its note-off.  State = (note in memory, frames since note-off), both float32 and
// updated with the reference's own arithmetic (products of saturated ramps, not branches):
//   activity    = min(relu(pitch), 1)
//   release_end = min(relu(steps - release_frames), 1)
//   out         = activity * pitch + (1 - activity) * previous * (1 - release_end)
//   steps       = (steps + 1) * (1 - activity) * (1 - release_end)
// One thread per voice row; the recurrence is sequential in time and a row is 750 frames.
__global__ void __launch_bounds__(128) note_release_kernel(const float* __restrict__ pitch,
                                                           float* __restrict__ out, int rows, int F,
                                                           int in_stride, float release_frames) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const float* p = pitch + (size_t)row * F * in_stride;
  float* o = out + (size_t)row * F;
  float previous = 0.f, steps = 0.f;
  for (int k = 0; k < F; ++k) {
    const float note = p[(size_t)k * in_stride];
    const float activity = fminf(fmaxf(note, 0.f), 1.f);
    const float release_end = fminf(fmaxf(__fadd_rn(steps, -release_frames), 0.f), 1.f);
    const float keep = __fmul_rn(__fmul_rn(__fadd_rn(1.f, -activity), previous),
                                 __fadd_rn(1.f, -release_end));
    const float y = __fadd_rn(__fmul_rn(activity, note), keep);
    steps = __fmul_rn(__fmul_rn(__fadd_rn(steps, 1.f), __fadd_rn(1.f, -activity)),
                      __fadd_rn(1.f, -release_end));
    previous = y;
    o[k] = y;
  }
}

}  // namespace b200ddsp
