// mel_sched_sim.cu - host-only model of the sparse-mel lane schedule (wwf_tables.h: build_mel_schedule): issue slots and
// shared-memory wavefronts (ideal vs with bank conflicts) per frame group, for a given feature configuration.
// build: nvcc -std=c++17 --expt-relaxed-constexpr -o tools/_bin/mel_sched_sim tools/mel_sched_sim.cu ; run: mel_sched_sim n_fft n_mels [pad]
#include <cstdio>
#include <cstdlib>
#include "../wakeword_trainer_home_b200/csrc/wwf_tables.h"
using namespace wwf;
int main(int argc, char** argv) {
  const int n_fft = argc > 1 ? atoi(argv[1]) : 400, M = argc > 2 ? atoi(argv[2]) : 40, pad = argc > 3 ? atoi(argv[3]) : 0;
  const int K = n_fft / 2 + 1;
  std::vector<float> fb = mel_fbanks32(K, 0.f, 8000.f, M, 16000);
  std::vector<int> lo(M), ofs(M + 1);
  std::vector<float> w;
  for (int m = 0; m < M; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < K; ++k) if (fb[(size_t)k * M + m] != 0.f) { if (first < 0) first = k; last = k; }
    ofs[m] = (int)w.size(); lo[m] = first < 0 ? 0 : first;
    if (first >= 0) for (int k = first; k <= last; ++k) w.push_back(fb[(size_t)k * M + m]);
  }
  ofs[M] = (int)w.size();
  (void)pad;
  auto zmap = [](int i) { return i; };                     // the power spectra are stored in plain bin order
  MelSchedule s = build_mel_schedule(lo, ofs, w, n_fft);
  long ideal = 0, actual = 0;
  for (int r = 0; r < s.rounds; ++r) {
    int mx = 0;
    for (int l = 0; l < 32; ++l) mx = std::max(mx, (int)((unsigned)s.tasks[r * 32 + l].x >> 16));
    for (int i = 0; i < mx; ++i)
      for (int h = 0; h < 2; ++h) {
        int cnt[16] = {0}, act = 0, mult = 0;
        for (int l = 16 * h; l < 16 * h + 16; ++l) {
          const int2 t = s.tasks[r * 32 + l];
          const int k0 = t.x & 0xffff, n = (int)((unsigned)t.x >> 16);
          if (i < n) { ++act; mult = std::max(mult, ++cnt[zmap(k0 + i) & 15]); }
        }
        if (act) { ideal += 1; actual += mult; }
      }
  }
  printf("n_fft %d mels %d: rounds %d, two-tap iterations %d, z-read wavefronts per frame pair: ideal %ld, with conflicts %ld (x%.2f)\n",
         n_fft, M, s.rounds, s.iterations, ideal, actual, (double)actual / ideal);
  return 0;
}
