// experiment: attempts per stop of every system (host build of the streamed solver)
#include <cstdint>
#include <cstring>
#include <math.h>
#include "../../chi_b200/csrc/integrator.cuh"
#include "../../oracle/csolver/port_models.inc"
using namespace chi_b200;
using M = Model_d8fb867e5f1e;
// out[b * max_stops + k] = attempts between stop k-1 and stop k
extern "C" int trace(const SolveArgs* a, int32_t* out, int max_stops) {
  for (int64_t b = 0; b < a->B; ++b) {
    double sm[stream_slots<M>()];
    StreamSolver<M, 11> s(*a, sm, 1);
    s.init(b);
    int k = 0, last = 0;
    double ts = -1;
    for (;;) {
      int ev = s.events();
      while (ev == s.EV_WAIT) { s.observe_parked(); ev = s.events(); }
      if (s.t_stop != ts) {
        if (ts >= 0 && k < max_stops) out[b * max_stops + k++] = s.n_att - last;
        last = s.n_att; ts = s.t_stop;
      }
      if (ev == s.EV_DONE || s.step()) break;
    }
    if (k < max_stops) out[b * max_stops + k++] = s.n_att - last;
    if (s.pend >= 0 && !s.invalid && s.status() == ST_OK) s.observe_parked();
    s.finish();
  }
  return 0;
}
