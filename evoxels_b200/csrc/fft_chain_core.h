// Work list of the chained z/y passes (fft_chain.cu), host/device.
//
// A "chain" runs two dependent passes over the x planes of the spectrum inside ONE persistent
// kernel so that the second pass finds the first one's output in L2 instead of HBM:
//   forward   stage 0 = z lines (real -> half spectrum) of a plane, stage 1 = y tiles of it
//   inverse   stage 0 = y tiles of a plane,                         stage 1 = z lines (+ u)
// Items are numbered in rounds.  Round k holds the n0 stage-0 items of plane k followed by the
// n1 stage-1 items of plane k - lag: a stage-1 item therefore comes `lag` rounds after the
// items it depends on.  Block b of a grid of G walks items b, b + G, b + 2G, ...; since every
// dependency points to a smaller item number and all blocks are co-resident (cooperative
// launch), waiting on a plane's completion counter cannot deadlock, and with a lag of a few
// grid-widths of items it practically never waits at all.
#pragma once
#include "evx_hd.h"

namespace evx {

struct ChainSchedule {
  int nplanes, lag, n0, n1;
  long long head;     // items of the first `lag` rounds (stage 0 only)
  long long body;     // items of the nplanes - lag full rounds
  long long total;
};

struct ChainItem {
  int stage, plane, idx;
};

EVX_HD ChainSchedule make_chain_schedule(int nplanes, int lag, int n0, int n1) {
  ChainSchedule s;
  s.nplanes = nplanes;
  s.lag = lag < 1 ? 1 : (lag > nplanes ? nplanes : lag);
  s.n0 = n0;
  s.n1 = n1;
  s.head = (long long)s.lag * n0;
  s.body = (long long)(nplanes - s.lag) * (n0 + n1);
  s.total = (long long)nplanes * (n0 + n1);
  return s;
}

EVX_HD ChainItem chain_decode(const ChainSchedule& s, long long i) {
  ChainItem it;
  if (i < s.head) {
    it.stage = 0;
    it.plane = (int)(i / s.n0);
    it.idx = (int)(i - (long long)it.plane * s.n0);
    return it;
  }
  long long j = i - s.head;
  if (j < s.body) {
    const int per = s.n0 + s.n1;
    const int round = (int)(j / per);
    const int w = (int)(j - (long long)round * per);
    if (w < s.n0) {
      it.stage = 0;
      it.plane = s.lag + round;
      it.idx = w;
    } else {
      it.stage = 1;
      it.plane = round;
      it.idx = w - s.n0;
    }
    return it;
  }
  j -= s.body;
  const int round = (int)(j / s.n1);
  it.stage = 1;
  it.plane = s.nplanes - s.lag + round;
  it.idx = (int)(j - (long long)round * s.n1);
  return it;
}

// ---- division-free form used by the kernel ----------------------------------------------
// "Virtual" numbering: nplanes + lag rounds of n0 + n1 slots each; slot w of round r is the
// stage-0 item w of plane r (if r < nplanes) or the stage-1 item w - n0 of plane r - lag (if
// r >= lag), else empty.  The same dependency rule holds (a stage-1 item sits `lag` rounds after
// its plane's stage-0 items), and a block steps from slot i to slot i + G with two additions.
struct ChainCursor {
  int round, w;            // current slot
  int step_round, step_w;  // G slots, decomposed
};
EVX_HD int chain_rounds(const ChainSchedule& s) { return s.nplanes + s.lag; }
EVX_HD void chain_cursor_init(ChainCursor& c, const ChainSchedule& s, int first, int stride) {
  const int per = s.n0 + s.n1;
  c.round = first / per;
  c.w = first - c.round * per;
  c.step_round = stride / per;
  c.step_w = stride - c.step_round * per;
}
EVX_HD void chain_cursor_step(ChainCursor& c, const ChainSchedule& s) {
  c.round += c.step_round;
  c.w += c.step_w;
  if (c.w >= s.n0 + s.n1) { c.w -= s.n0 + s.n1; ++c.round; }
}
EVX_HD bool chain_cursor_done(const ChainCursor& c, const ChainSchedule& s) { return c.round >= chain_rounds(s); }
// item at the cursor; false for an empty slot
EVX_HD bool chain_cursor_item(const ChainCursor& c, const ChainSchedule& s, ChainItem& it) {
  if (c.w < s.n0) {
    it.stage = 0; it.plane = c.round; it.idx = c.w;
    return c.round < s.nplanes;
  }
  it.stage = 1; it.plane = c.round - s.lag; it.idx = c.w - s.n0;
  return c.round >= s.lag;
}
// advance to the next non-empty slot at or after the cursor; false when the list is exhausted
EVX_HD bool chain_cursor_next(ChainCursor& c, const ChainSchedule& s, ChainItem& it) {
  while (!chain_cursor_done(c, s)) {
    if (chain_cursor_item(c, s, it)) return true;
    chain_cursor_step(c, s);
  }
  return false;
}

}  // namespace evx
