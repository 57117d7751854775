// Host-side voice allocation: pianoroll -> polyphonic conditioning (reference
// ddsp_piano/utils/midi_encoders.py:4-104, MIDIRoll2Conditioning).  Pure host
// code (a stateful scan over frames); included by b200ddsp.cu so that it ships in the same library.
//
// Per frame the reference keeps the n_synths highest active pitches, then permutes them into
// channels so that a sounding note stays on its channel: sustained notes keep their slot, new
// notes take the next free slot of a round-robin `assigner`, silent entries fill what is left.
// The port keeps the reference's behaviour to the letter, including its corner cases: the
// assigner is -1 while every channel is busy and is then used as a Python index (= last
// channel), and a frame whose pitch SET equals the assigned set reuses the previous permutation.
#include <algorithm>

namespace {

struct VoiceAllocator {
  int n;
  int assigner = 0;
  std::vector<float> assigned;   // pitch held by each channel, 0 = free
  std::vector<int> reorder;
  explicit VoiceAllocator(int n_) : n(n_), assigned(n_, 0.f), reorder(n_) {
    for (int i = 0; i < n; ++i) reorder[i] = i;
  }
  bool is_assigned(float p) const { return std::find(assigned.begin(), assigned.end(), p) != assigned.end(); }
  int slot() const { return assigner < 0 ? n + assigner : assigner; }   // Python negative index
  void update_assigner() {                                              // midi_encoders.py:22-31
    assigner = (assigner + 1) % n;
    if (!is_assigned(0.f)) {
      assigner = -1;
    } else {
      while (assigned[assigner] != 0.f) assigner = (assigner + 1) % n;
    }
  }
};

std::vector<float> sorted_unique(const float* v, int n) {
  std::vector<float> u(v, v + n);
  std::sort(u.begin(), u.end());
  u.erase(std::unique(u.begin(), u.end()), u.end());
  return u;
}

}  // namespace

extern "C" int b200ddsp_midi_roll_to_conditioning(const float* roll, int n_frames, int n_pitches,
                                                  int n_synths, float first_pitch,
                                                  float* conditioning, float* polyphony) {
  if (!roll || !conditioning || n_frames < 0 || n_pitches < 1 || n_synths < 1 || n_synths > n_pitches)
    return B200DDSP_BAD_ARGUMENT;
  VoiceAllocator va(n_synths);
  const int n = n_synths;
  std::vector<std::pair<float, int>> keyed(n_pitches);
  std::vector<float> pitches(n), vel(n);
  for (int t = 0; t < n_frames; ++t) {
    const float* fr = roll + (size_t)t * n_pitches * 2;
    // polyphony = number of active notes; activity * pitch number   (:46-50)
    float poly = 0.f;
    for (int i = 0; i < n_pitches; ++i) {
      poly += fr[2 * i];
      keyed[i] = {fr[2 * i] * (first_pitch + (float)i), i};
    }
    if (polyphony) polyphony[t] = poly;
    // the n highest entries in ascending order (np.argsort(...)[:, -n:]); ties among silent
    // entries are broken by index (stable), the reference leaves them to the sort
    std::stable_sort(keyed.begin(), keyed.end(),
                     [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first < b.first; });
    for (int c = 0; c < n; ++c) {
      const auto& e = keyed[n_pitches - n + c];
      pitches[c] = e.first;
      vel[c] = fr[2 * e.second + 1];
    }
    // same pitch set as the channels hold: reuse the permutation   (:60-68)
    const std::vector<float> up = sorted_unique(pitches.data(), n);
    const std::vector<float> ua = sorted_unique(va.assigned.data(), n);
    std::vector<float> common;
    std::set_intersection(up.begin(), up.end(), ua.begin(), ua.end(), std::back_inserter(common));
    const bool unchanged = t > 0 && common.size() == up.size() && common.size() == ua.size();
    if (!unchanged) {
      std::vector<int> reorder(n, 0);
      auto in_frame = [&](float p) { return std::find(pitches.begin(), pitches.end(), p) != pitches.end(); };
      // free the channels whose note has ended   (:72-78)
      for (int c = 0; c < n; ++c) {
        if (!in_frame(va.assigned[c])) {
          va.assigned[c] = 0.f;
          if (va.assigner == -1) va.update_assigner();
        }
      }
      // sustained notes stay where they are   (:80-84)
      for (int c = 0; c < n; ++c) {
        if (pitches[c] != 0.f && va.is_assigned(pitches[c])) {
          const int ch = (int)(std::find(va.assigned.begin(), va.assigned.end(), pitches[c]) - va.assigned.begin());
          reorder[ch] = c;
        }
      }
      // new notes take the next free channel   (:86-91)
      for (int c = 0; c < n; ++c) {
        if (!va.is_assigned(pitches[c])) {
          reorder[va.slot()] = c;
          va.assigned[va.slot()] = pitches[c];
          va.update_assigner();
        }
      }
      // silent entries fill the remaining channels   (:93-97)
      for (int c = 0; c < n; ++c) {
        if (pitches[c] == 0.f) {
          reorder[va.slot()] = c;
          va.update_assigner();
        }
      }
      va.reorder = reorder;
    }
    float* out = conditioning + (size_t)t * n * 2;
    for (int ch = 0; ch < n; ++ch) {
      out[2 * ch] = pitches[va.reorder[ch]];
      out[2 * ch + 1] = vel[va.reorder[ch]];
    }
  }
  return B200DDSP_OK;
}
