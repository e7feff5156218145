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

}  // namespace evx
