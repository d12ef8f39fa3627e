// Corintho game rules on a bit-plane packed state -- branch-light SWAR code shared by the
// game-logic kernel (one thread per state) and the tree kernels (warp-uniform expand).
//
// Packed state ("cstate", 16 bytes; include/corintho_b200.h):
//   w0: four 16-bit planes, bit s = row*4+col of plane t at bit 16*t+s,
//       t = 0 base, 1 column, 2 capital, 3 frozen
//       (reference: Game::board_ bit row*16+col*4+t, cpp/include/game.h:124-126)
//   w1: byte i<6 = pieces_[i] (P0{B,C,A},P1{B,C,A}, game.h:127-131), byte 6 = to_play_
//
// Behaviour restated from the reference (never its code):
//   legal moves   cpp/src/game.cpp:28-43 (getLegalMoves), 193-242 (canPlace/canMove),
//                 249-405 (line rules, first-found line per category, capital fix-ups)
//   move codec    cpp/src/move.cpp:11-42
//   do_move       cpp/src/game.cpp:60-96
//   NN encoding   cpp/src/game.cpp:45-58
//   terminal test cpp/src/node.cpp:256-271
#ifndef CORINTHO_B200_RULES_CUH
#define CORINTHO_B200_RULES_CUH

#include <stdint.h>

#if defined(__CUDACC__)
#define CB_HD __host__ __device__ __forceinline__
#else
#define CB_HD inline
#endif

namespace cb200 {

struct CState {
  uint64_t w0, w1;
};

enum : int {
  kResultNone = 0,
  kResultLoss = 1,
  kResultDraw = 2,
  kResultWin = 3,
  kDeducedLoss = 4,
  kDeducedDraw = 5,
  kDeducedWin = 6
};

CB_HD int cb_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
CB_HD int cb_ffs(uint32_t x) {  // 1-based index of lowest set bit, 0 if none
#if defined(__CUDA_ARCH__)
  return __ffs((int)x);
#else
  return __builtin_ffs((int)x);
#endif
}

CB_HD CState start_state() {
  CState s;
  s.w0 = 0;
  s.w1 = 0x0000040404040404ull;
  return s;
}

// 3-in-4 row index compression: bits {4r+c, c<3} -> {3r+c}
CB_HD uint32_t compress3(uint32_t v) {
  return (v & 0x7u) | ((v >> 1) & 0x38u) | ((v >> 2) & 0x1C0u) | ((v >> 3) & 0xE00u);
}

// AND the 96-bit line-breaker mask `idx` into (m0, m1, m2). Device tables are padded to four
// 16-byte-aligned words, so one 128-bit load fetches a mask.
template <class LBFn>
CB_HD void lb_and(LBFn LB, int idx, uint32_t &m0, uint32_t &m1, uint32_t &m2) {
  const uint32_t *lb = LB(idx);
#ifdef __CUDA_ARCH__
  const uint4 v = *reinterpret_cast<const uint4 *>(lb);
  m0 &= v.x, m1 &= v.y, m2 &= v.z;
#else
  m0 &= lb[0], m1 &= lb[1], m2 &= lb[2];
#endif
}

// Legal-move mask (96 bits in m[0..2], bit id&31 of word id>>5) and the "lines present" flag.
// LB(idx) must return a pointer to the 3 words of line-breaker mask idx (util.h:85-290 data);
// idx 102 must return an all-ones mask ("no line in this category").
// kBranchless: thread-per-state callers (K1) get select-only code (every category ANDs a table
// entry, 102 when there is no line); warp-uniform callers keep early-outs that skip work.
template <bool kBranchless, class LBFn>
CB_HD bool legal_moves_t(const CState &st, uint32_t m[3], LBFn LB) {
  const uint32_t lo = (uint32_t)st.w0, hi = (uint32_t)(st.w0 >> 32);
  const uint32_t B = lo & 0xFFFFu, C = lo >> 16, A = hi & 0xFFFFu, F = hi >> 16;
  const uint32_t O = B | C | A;
  const uint32_t E = ~O & 0xFFFFu;
  // squares by top piece (game.cpp:158-168) and by bottom piece (170-180)
  const uint32_t T2 = A, T1 = C & ~A, T0 = B & ~(C | A);
  const uint32_t bot1 = C & ~B, bot2 = A & ~(B | C);
  const uint32_t nF = ~F;
  // canMove (game.cpp:222-232): both non-empty, neither frozen, bottom(from)-top(to)==1
  const uint32_t X1 = bot1 & nF, X2 = bot2 & nF, Y0 = T0 & nF, Y1 = T1 & nF;
  const uint32_t mr = ((X1 & (Y0 >> 1)) | (X2 & (Y1 >> 1))) & 0x7777u;
  const uint32_t md = ((X1 & (Y0 >> 4)) | (X2 & (Y1 >> 4))) & 0x0FFFu;
  const uint32_t ml = ((X1 & (Y0 << 1)) | (X2 & (Y1 << 1))) & 0xEEEEu;
  const uint32_t mu = ((X1 & (Y0 << 4)) | (X2 & (Y1 << 4))) & 0xFFF0u;
  const uint32_t R12 = compress3(mr), L12 = compress3(ml >> 1), D12 = md, U12 = mu >> 4;
  // canPlace (game.cpp:193-220)
  const uint32_t tp = (uint32_t)(st.w1 >> 48) & 1u;
  const uint32_t pcs = (uint32_t)(st.w1 >> (24 * tp)) & 0xFFFFFFu;
  const uint32_t pb = (pcs & 0xFFu) ? E : 0u;
  const uint32_t pc = (pcs & 0xFF00u) ? (E | Y0) : 0u;
  const uint32_t pa = (pcs & 0xFF0000u) ? (E | Y1) : 0u;
  uint32_t m0 = R12 | (D12 << 12) | (L12 << 24);
  uint32_t m1 = (L12 >> 8) | (U12 << 4) | (pb << 16);
  uint32_t m2 = pc | (pa << 16);
  bool lines = false;

  // ---- row lines (game.cpp:249-315, isCol=false): first row with a line, long > {0,1,2} > {1,2,3}
  {
    const uint32_t W0 = T0 & (T0 >> 1) & (T0 >> 2), W1 = T1 & (T1 >> 1) & (T1 >> 2),
                   W2 = T2 & (T2 >> 1) & (T2 >> 2);
    const uint32_t any1 = (W1 | (W1 >> 1)) & 0x1111u, any2 = (W2 | (W2 >> 1)) & 0x1111u;
    const uint32_t Wa = W0 | W1 | W2;
    const uint32_t left = Wa & 0x1111u, right = (Wa >> 1) & 0x1111u;
    const uint32_t any = left | right;
    if (kBranchless || any) {
      const bool has = any != 0;
      lines |= has;
      const int i4 = has ? cb_ffs(any) - 1 : 0, i = i4 >> 2;
      const int t = (int)((any1 >> i4) & 1u) + 2 * (int)((any2 >> i4) & 1u);
      const bool l = (left >> i4) & 1u, r = (right >> i4) & 1u;
      const int cat = (l && r) ? 2 : (l ? 0 : 1);  // RB, RL, RR (util.h:67-69)
      lb_and(LB, has ? cat * 12 + i * 3 + t : 102, m0, m1, m2);
      if (has && t == 2 && cat != 2) {  // capital fix-ups (game.cpp:280-309), column `e`
        const int e = l ? 3 : 0;
        const uint32_t colm = 0x111u << e;
        m0 &= ~(colm << 12) | ((A & colm) << 12);        // down moves from (k,e), k=0..2
        m1 &= ~(colm << 4) | (((A >> 4) & colm) << 4);   // up moves from (k,e), k=1..3
      }
    }
  }
  // ---- column lines (isCol=true): first column, long > rows{0,1,2} > rows{1,2,3}
  {
    const uint32_t W0 = T0 & (T0 >> 4) & (T0 >> 8), W1 = T1 & (T1 >> 4) & (T1 >> 8),
                   W2 = T2 & (T2 >> 4) & (T2 >> 8);
    const uint32_t any1 = (W1 | (W1 >> 4)) & 0xFu, any2 = (W2 | (W2 >> 4)) & 0xFu;
    const uint32_t Wa = W0 | W1 | W2;
    const uint32_t upper = Wa & 0xFu, lower = (Wa >> 4) & 0xFu;
    const uint32_t any = upper | lower;
    if (kBranchless || any) {
      const bool has = any != 0;
      lines |= has;
      const int i = has ? cb_ffs(any) - 1 : 0;
      const int t = (int)((any1 >> i) & 1u) + 2 * (int)((any2 >> i) & 1u);
      const bool u = (upper >> i) & 1u, d = (lower >> i) & 1u;
      const int cat = (u && d) ? 5 : (u ? 3 : 4);  // CB, CU, CD (util.h:70-72)
      lb_and(LB, has ? cat * 12 + i * 3 + t : 102, m0, m1, m2);
      if (has && t == 2 && cat != 5) {  // capital fix-ups along row `e`
        const int e = u ? 3 : 0;
        const uint32_t Arow = (A >> (4 * e)) & 0xFu;
        const uint32_t rowm = 7u << (3 * e);
        const uint32_t keepR = ~rowm | ((Arow & 7u) << (3 * e));         // right moves from (e,k)
        const uint32_t keepL = ~rowm | (((Arow >> 1) & 7u) << (3 * e));  // left moves from (e,k)
        m0 &= (keepR | ~0xFFFu) & ((keepL << 24) | 0x00FFFFFFu);
        m1 &= ((keepL & 0xFFFu) >> 8) | ~0xFu;
      }
    }
  }
  // ---- diagonals. Three equal tops with square step 5 (bases 0,5 = main diagonal upper/lower,
  // 1 = S1, 4 = S3) or step 3 (bases 3,6 = anti-diagonal upper/lower, 2 = S0, 7 = S2).
  {
    const uint32_t A5 = T0 & (T0 >> 5) & (T0 >> 10), B5 = T1 & (T1 >> 5) & (T1 >> 10),
                   C5 = T2 & (T2 >> 5) & (T2 >> 10);
    const uint32_t A3 = T0 & (T0 >> 3) & (T0 >> 6), B3 = T1 & (T1 >> 3) & (T1 >> 6),
                   C3 = T2 & (T2 >> 3) & (T2 >> 6);
    const uint32_t D5 = (A5 | B5 | C5) & 0x33u, D3 = (A3 | B3 | C3) & 0xCCu;
    if (kBranchless || (D5 | D3)) {
      // long diagonals (game.cpp:317-360): main then anti; long > upper > lower
      const bool mu5 = D5 & 0x01u, ml5 = D5 & 0x20u, au3 = D3 & 0x08u, al3 = D3 & 0x40u;
      const bool mainl = mu5 || ml5, antil = au3 || al3;
      int D = mainl ? ((mu5 && ml5) ? 2 : (mu5 ? 0 : 1))    // D0B, D0U, D0D
                    : ((au3 && al3) ? 5 : (au3 ? 3 : 4));   // D1B, D1U, D1D
      int base = mainl ? (mu5 ? 0 : 5) : (au3 ? 3 : 6);
      uint32_t b1 = mainl ? B5 : B3, b2 = mainl ? C5 : C3;
      int t = (int)((b1 >> base) & 1u) + 2 * (int)((b2 >> base) & 1u);
      bool has = mainl || antil;
      lines |= has;
      lb_and(LB, has ? 72 + D * 3 + t : 102, m0, m1, m2);
      // short diagonals (game.cpp:362-391): S0..S3, first found
      const bool s0 = D3 & 0x04u, s1 = D5 & 0x02u, s2 = D3 & 0x80u, s3 = D5 & 0x10u;
      D = s0 ? 6 : (s1 ? 7 : (s2 ? 8 : 9));
      base = s0 ? 2 : (s1 ? 1 : (s2 ? 7 : 4));
      const bool st5 = !s0 && (s1 || (!s2 && s3));
      b1 = st5 ? B5 : B3, b2 = st5 ? C5 : C3;
      t = (int)((b1 >> base) & 1u) + 2 * (int)((b2 >> base) & 1u);
      has = s0 || s1 || s2 || s3;
      lines |= has;
      lb_and(LB, has ? 72 + D * 3 + t : 102, m0, m1, m2);
    }
  }
  m[0] = m0, m[1] = m1, m[2] = m2;
  return lines;
}
// First half of legal_moves_t for callers that treat positions with a line separately (K1): the
// basic-rule mask (canPlace / canMove only) and whether ANY line exists on the board -- the
// union of the four `has` conditions above. Without a line the basic mask IS the legal mask.
CB_HD bool basic_moves(const CState &st, uint32_t m[3]) {
  const uint32_t lo = (uint32_t)st.w0, hi = (uint32_t)(st.w0 >> 32);
  const uint32_t B = lo & 0xFFFFu, C = lo >> 16, A = hi & 0xFFFFu, F = hi >> 16;
  const uint32_t O = B | C | A;
  const uint32_t E = ~O & 0xFFFFu;
  const uint32_t T2 = A, T1 = C & ~A, T0 = B & ~(C | A);
  const uint32_t bot1 = C & ~B, bot2 = A & ~(B | C);
  const uint32_t nF = ~F;
  const uint32_t X1 = bot1 & nF, X2 = bot2 & nF, Y0 = T0 & nF, Y1 = T1 & nF;
  const uint32_t mr = ((X1 & (Y0 >> 1)) | (X2 & (Y1 >> 1))) & 0x7777u;
  const uint32_t md = ((X1 & (Y0 >> 4)) | (X2 & (Y1 >> 4))) & 0x0FFFu;
  const uint32_t ml = ((X1 & (Y0 << 1)) | (X2 & (Y1 << 1))) & 0xEEEEu;
  const uint32_t mu = ((X1 & (Y0 << 4)) | (X2 & (Y1 << 4))) & 0xFFF0u;
  const uint32_t R12 = compress3(mr), L12 = compress3(ml >> 1), D12 = md, U12 = mu >> 4;
  const uint32_t tp = (uint32_t)(st.w1 >> 48) & 1u;
  const uint32_t pcs = (uint32_t)(st.w1 >> (24 * tp)) & 0xFFFFFFu;
  const uint32_t pb = (pcs & 0xFFu) ? E : 0u;
  const uint32_t pc = (pcs & 0xFF00u) ? (E | Y0) : 0u;
  const uint32_t pa = (pcs & 0xFF0000u) ? (E | Y1) : 0u;
  m[0] = R12 | (D12 << 12) | (L12 << 24);
  m[1] = (L12 >> 8) | (U12 << 4) | (pb << 16);
  m[2] = pc | (pa << 16);
  // three equal tops in a row (step 1, starts in columns 0-1), in a column (step 4, rows 0-1), on
  // a diagonal (step 5 from squares 0, 1, 4, 5; step 3 from squares 2, 3, 6, 7): the three planes
  // are disjoint, so each step needs one AND chain per plane
  uint32_t row = 0, col = 0, d5 = 0, d3 = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int p = 0; p < 3; ++p) {
    const uint32_t T = p == 0 ? T0 : (p == 1 ? T1 : T2);
    row |= T & (T >> 1) & (T >> 2);
    col |= T & (T >> 4) & (T >> 8);
    d5 |= T & (T >> 5) & (T >> 10);
    d3 |= T & (T >> 3) & (T >> 6);
  }
  return ((row & 0x3333u) | (col & 0x00FFu) | (d5 & 0x33u) | (d3 & 0xCCu)) != 0u;
}

template <class LBFn>
CB_HD bool legal_moves(const CState &st, uint32_t m[3], LBFn LB) {
  return legal_moves_t<false>(st, m, LB);
}

// game.cpp:60-96 (no legality check, like the reference); select-only code
CB_HD CState do_move(const CState &st, int move) {
  CState o;
  uint64_t w0 = st.w0 & 0x0000FFFFFFFFFFFFull;  // clear every frozen bit
  uint64_t w1 = st.w1;
  const uint32_t tp = (uint32_t)(w1 >> 48) & 1u;
  const bool is_place = move >= 48;
  // place: 48 + piece*16 + square
  const int piece = (move - 48) >> 4;
  // move: dir = id/12 (0 right, 1 down, 2 left, 3 up), r = id%12 (move.cpp:11-42)
  // (small unsigned multiply-shift divisions: exact for ids 0..47 / r 0..11)
  const uint32_t um = (uint32_t)move & 63u;
  const int dir = (int)((um * 43u) >> 9), r = (int)um - dir * 12;
  const int rq = (int)(((uint32_t)r * 11u) >> 5);  // r / 3
  const int r3 = rq * 4 + (r - rq * 3);            // (row, col<3) of a horizontal move
  const int from = dir == 0 ? r3 : (dir == 1 ? r : (dir == 2 ? r3 + 1 : r + 4));
  const int to_m = dir == 0 ? r3 + 1 : (dir == 1 ? r + 4 : (dir == 2 ? r3 : r));
  const int to = is_place ? (move & 15) : to_m;
  const int fsh = is_place ? 0 : from;
  const uint64_t planes = is_place ? 0ull : 0x0000000100010001ull;
  const uint64_t stack = (w0 >> fsh) & planes;
  w0 &= ~(planes << fsh);
  w0 |= stack << to;
  w0 |= is_place ? (1ull << (16 * (piece & 3) + to)) : 0ull;
  w0 |= 1ull << (48 + to);
  w1 -= is_place ? (1ull << (8 * (tp * 3 + (uint32_t)(piece & 3)))) : 0ull;
  w1 ^= 1ull << 48;
  o.w0 = w0, o.w1 = w1;
  return o;
}

// One element of the 70-float NN input (game.cpp:45-58)
CB_HD float encode_elem(const CState &st, int j) {
  if (j < 64) return (float)((st.w0 >> (16 * (j & 3) + (j >> 2))) & 1ull);
  const uint32_t tp = (uint32_t)(st.w1 >> 48) & 1u;
  int i = (int)tp * 3 + (j - 64);
  if (i >= 6) i -= 6;
  return (float)((st.w1 >> (8 * i)) & 0xFFull) * 0.25f;
}

// node.cpp:256-271: no legal move -> the mover lost if a line exists, else draw
CB_HD int terminal_result(int n_legal, bool lines) {
  return n_legal == 0 ? (lines ? kResultLoss : kResultDraw) : kResultNone;
}

// id of the k-th (0-based) set bit of the 96-bit mask (popcount binary search, no loops)
CB_HD int nth_move(const uint32_t m[3], int k) {
  const int c0 = cb_popc(m[0]), c1 = cb_popc(m[1]);
  const bool in0 = k < c0, in1 = k < c0 + c1;
  uint32_t w = in0 ? m[0] : (in1 ? m[1] : m[2]);
  int base = in0 ? 0 : (in1 ? 32 : 64);
  k -= in0 ? 0 : (in1 ? c0 : c0 + c1);
#pragma unroll
  for (int sh = 16; sh >= 1; sh >>= 1) {
    const int c = cb_popc(w & ((1u << sh) - 1u));
    const bool up = k >= c;
    k -= up ? c : 0;
    w = up ? (w >> sh) : w;
    base += up ? sh : 0;
  }
  return base;
}

// deterministic per-state random word of the game-logic workload (splitmix64 finaliser)
CB_HD uint32_t step_rnd(uint64_t seed, uint64_t i) {
  uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}

}  // namespace cb200
#endif
