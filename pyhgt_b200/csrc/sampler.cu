// Host-side helper of pyhgt_b200/sampler.py (HGSampling, reference pyHGT/data.py:87-210).  Plain C++ (no device code):
// the budget update of ONE adjacency slice — data.py:123-130, the loop over the sampled neighbours of one
// <target type, source type, relation> block — done in one call instead of a dozen small numpy operations.
// The random draws stay in numpy (same global RNG stream as the reference); this function is deterministic.
#include "common.cuh"

// ids / tms: the sampled neighbours and their edge times (no_time marks the reference's `None`), in sampling order.
// in_layer [n]: membership of layer_data[source_type];  in_budget / score / b_time / stamp [n]: budget[source_type] as
// flat arrays (stamp reproduces the dict's insertion order);  *stamp_counter: next insertion stamp (shared by all types).
// *touched_layer is set when some candidate passes the time test (the reference then evaluates
// `source_id in layer_data[source_type]`, which creates that defaultdict entry).
// Returns the number of candidates added / updated, or -1 when an id is outside [0, n).
extern "C" int64_t hgt_sampler_budget_update(const int64_t* ids, const int64_t* tms, int64_t n_s, int64_t target_time,
                                             int64_t no_time, int64_t max_time, int64_t n, const uint8_t* in_layer,
                                             uint8_t* in_budget, double* score, int64_t* b_time, int64_t* stamp,
                                             int64_t* stamp_counter, int32_t* touched_layer) {
  for (int64_t i = 0; i < n_s; ++i)                         // validate before touching anything: -1 leaves the state intact
    if (ids[i] < 0 || ids[i] >= n) return -1;
  int64_t kept = 0;
  const double w = 1.0 / (double)n_s;                       // 1. / len(sampled_ids), data.py:129
  for (int64_t i = 0; i < n_s; ++i) {
    const int64_t tm = tms[i] == no_time ? target_time : tms[i];
    if (tm > max_time) continue;                            // data.py:127, first operand of the `or`
    *touched_layer = 1;
    const int64_t id = ids[i];
    if (in_layer[id]) continue;                             // second operand
    if (!in_budget[id]) {                                   // defaultdict creates [0., 0] at the end of the dict
      in_budget[id] = 1;
      score[id] = 0.0;
      stamp[id] = (*stamp_counter)++;
    }
    score[id] += w;
    b_time[id] = tm;
    ++kept;
  }
  return kept;
}
