// TEST INFRASTRUCTURE: compiles the product's SWAR rules header (corintho_ai_b200/csrc/rules.cuh)
// for the HOST so its logic can be checked against the oracle on machines without a GPU.
// This is not a CPU fallback: nothing in the product links it.
#include <stdint.h>
#include "../../corintho_ai_b200/csrc/corintho_tables.h"
#include "../../corintho_ai_b200/csrc/rules.cuh"

extern "C" void shim_step_batch(int64_t n, const uint64_t *states, uint64_t seed, uint32_t *maskflags,
                                uint64_t *next, float *enc) {
  using namespace cb200;
  static const uint32_t ones[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
  auto LB = [](int idx) { return idx < 102 ? kCLineBreakers[idx] : ones; };
  for (int64_t i = 0; i < n; ++i) {
    CState s{states[2 * i], states[2 * i + 1]};
    uint32_t m[3];
    bool lines = (i & 1) ? legal_moves_t<true>(s, m, LB) : legal_moves_t<false>(s, m, LB);
    int nl = cb_popc(m[0]) + cb_popc(m[1]) + cb_popc(m[2]);
    int result = terminal_result(nl, lines);
    int chosen = 0x7f;
    CState o = s;
    if (nl > 0) {
      chosen = nth_move(m, (int)(step_rnd(seed, (uint64_t)i) % (uint32_t)nl));
      o = do_move(s, chosen);
    }
    maskflags[4 * i] = m[0], maskflags[4 * i + 1] = m[1], maskflags[4 * i + 2] = m[2];
    maskflags[4 * i + 3] = (uint32_t)result | (lines ? 4u : 0u) | ((uint32_t)nl << 8) | ((uint32_t)chosen << 16);
    next[2 * i] = o.w0, next[2 * i + 1] = o.w1;
    if (enc) for (int j = 0; j < 70; ++j) enc[70 * i + j] = encode_elem(s, j);
  }
}
extern "C" uint32_t shim_step_rnd(uint64_t seed, uint64_t i) { return cb200::step_rnd(seed, i); }

// basic_moves() (the first half of legal_moves_t that the split K1 kernel runs on every position):
// its "any line" flag must equal legal_moves_t's, and without a line its mask IS the legal mask.
// Returns the number of positions where that fails.
extern "C" int64_t shim_basic_moves_check(int64_t n, const uint64_t *states) {
  using namespace cb200;
  static const uint32_t ones[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
  auto LB = [](int idx) { return idx < 102 ? kCLineBreakers[idx] : ones; };
  int64_t bad = 0;
  for (int64_t i = 0; i < n; ++i) {
    CState s{states[2 * i], states[2 * i + 1]};
    uint32_t m[3], b[3];
    const bool lines = legal_moves_t<false>(s, m, LB);
    const bool any = basic_moves(s, b);
    if (any != lines) ++bad;
    else if (!any && (b[0] != m[0] || b[1] != m[1] || b[2] != m[2])) ++bad;
    else if (any && ((m[0] & ~b[0]) | (m[1] & ~b[1]) | (m[2] & ~b[2]))) ++bad;  // line rules only remove moves
  }
  return bad;
}
