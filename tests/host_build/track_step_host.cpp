// TEST-ONLY host build of the tracker-step routine: the very source the CUDA kernel is compiled from
// (trackformer_b200/csrc/track_step_core.h) run by one "thread", so the CPU suite can replay the reference-recorded
// tracking sequences through the device algorithm.  Built by tests/test_track_step_cpu.py; never shipped, never loaded
// by the package.
#include "../../trackformer_b200/csrc/track_step_core.h"

extern "C" int tfb200_track_step_host(const TfbTrackStepArgs* a) {
  static tfb200_track::Shared sh;
  tfb200_track::ts_step(*a, sh);
  return 0;
}
