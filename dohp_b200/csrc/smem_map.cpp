// Host side of the conflict-free shared-memory scheme of eo_stages.cuh: for one kernel configuration (n quadrature points per
// direction, EPB elements per block, base element stride) choose the padding of the element stride and deal the x lines (b,c) and the
// y lines (a,c) of a block to threads so that the 16 lanes of every half-warp touch 16 different 64-bit banks at every step of a pass.
// A pass is "thread owns a line, steps along it": the bank of a line at step m is (slot*ES + u*S1 + v*S2 + m*S3) mod 16, so a deal that
// is conflict free at m = 0 is conflict free at every step.  The z lines (a,b) keep the natural thread order (their threads also
// address the stash planes in HBM, which must stay coalesced); the padding is chosen so that this order is conflict free as well.
// Setup code, runs once per (context, configuration); the tables go to device memory.
#include "internal.h"
#include <algorithm>
#include <map>
#include <tuple>

namespace dohp {

static int natural_wavefronts(const std::vector<int> &bank)
{
  int tot = 0;
  for (size_t h0 = 0; h0 < bank.size(); h0 += 16) {
    int cnt[16] = {0}, mx = 0;
    for (size_t i = h0; i < std::min(bank.size(), h0 + 16); i++) mx = std::max(mx, ++cnt[bank[i]]);
    tot += mx;
  }
  return tot;
}

// deal `bank.size()` lines to half-warps of 16 lanes (the last one partial): returns the wavefronts per step and the line of every thread
static int deal(const std::vector<int> &bank, std::vector<int> &line_of_thread)
{
  const int L = (int)bank.size(), H = (L + 15) / 16;
  std::vector<int> cap(H, 16);
  cap[H - 1] = L - 16 * (H - 1);
  std::vector<std::vector<int>> hw(H);
  std::vector<std::array<int, 16>> mult(H);
  for (auto &m : mult) m.fill(0);
  std::vector<int> byb[16];
  for (int i = 0; i < L; i++) byb[bank[i]].push_back(i);
  int order[16];
  for (int b = 0; b < 16; b++) order[b] = b;
  std::stable_sort(order, order + 16, [&](int x, int y) { return byb[x].size() > byb[y].size(); });
  std::vector<std::pair<int, int>> left;
  for (int ob = 0; ob < 16; ob++) {
    const int b = order[ob];
    std::vector<int> ids = byb[b];
    std::vector<int> idx(H);
    for (int h = 0; h < H; h++) idx[h] = h;
    std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return cap[x] - (int)hw[x].size() > cap[y] - (int)hw[y].size(); });
    for (int h : idx) {
      if (ids.empty()) break;
      if (cap[h] - (int)hw[h].size() > 0) { hw[h].push_back(ids.back()); ids.pop_back(); mult[h][b]++; }
    }
    for (int i : ids) left.push_back({i, b});
  }
  for (auto &lb : left) {
    int best = -1;
    for (int h = 0; h < H; h++)
      if (cap[h] - (int)hw[h].size() > 0 && (best < 0 || mult[h][lb.second] < mult[best][lb.second])) best = h;
    hw[best].push_back(lb.first);
    mult[best][lb.second]++;
  }
  int wf = 0;
  line_of_thread.assign(L, -1);
  for (int h = 0; h < H; h++) {
    std::sort(hw[h].begin(), hw[h].end());                 // neighbouring lanes keep neighbouring lines (global gathers of the x pass)
    for (size_t k = 0; k < hw[h].size(); k++) line_of_thread[h * 16 + k] = hw[h][k];
    wf += *std::max_element(mult[h].begin(), mult[h].end());
  }
  return wf;
}

// nq: points per direction; sa, sb: layout strides (Lay2); epb: elements per block; threads: block size; es_base: element stride
// before padding.  Returns the padded element stride and the maps [2][threads] (x pass, y pass), entry = slot<<16 | u<<8 | v.
int build_smem_maps(int nq, int sa, int sb, int epb, int threads, int es_base, int *es_out, std::vector<int> &maps)
{
  const int T = nq * nq, L = epb * T;
  if (L > threads) return seterr(77, "smem map: %d lines for %d threads", L, threads);
  auto banks = [&](int es, int s1, int s2) {
    std::vector<int> b(L);
    for (int le = 0; le < epb; le++)
      for (int u = 0; u < nq; u++)
        for (int v = 0; v < nq; v++) b[le * T + u * nq + v] = (le * es + u * s1 + v * s2) & 15;
    return b;
  };
  long best = -1;
  int best_pad = 0;
  std::vector<int> lx, ly, bx, by;
  for (int pad = 0; pad < 16; pad++) {
    const int es = es_base + pad;
    std::vector<int> tx, ty;
    // weights: accesses per pass of the scalar kernel (z: 9 n^3, x: 6 n^3, y: 8 n^3)
    const long tot = 9L * natural_wavefronts(banks(es, sa, sb)) + 6L * deal(banks(es, sb, 1), tx) + 8L * deal(banks(es, sa, 1), ty);
    if (best < 0 || tot < best) { best = tot; best_pad = pad; bx = tx; by = ty; }
    if (epb == 1) break;                                   // the stride between elements does not matter
  }
  *es_out = es_base + best_pad;
  maps.assign((size_t)2 * threads, epb << 16);             // slot = epb: no line
  for (int pass = 0; pass < 2; pass++) {
    const std::vector<int> &lt = pass ? by : bx;
    std::vector<char> seen(L, 0);
    for (int t = 0; t < L; t++) {
      const int line = lt[t];
      if (line < 0 || line >= L || seen[line]) return seterr(77, "smem map: not a bijection");
      seen[line] = 1;
      const int le = line / T, r = line - le * T;
      maps[(size_t)pass * threads + t] = le << 16 | (r / nq) << 8 | (r % nq);
    }
  }
  return 0;
}

}  // namespace dohp
