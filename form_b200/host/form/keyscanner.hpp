// Key-scan bookkeeping: which scans stay in the fixed-lag window.
//
// Host-side, O(60) integer work per scan; mirrors the behaviour of
// form::KeyScanner (/root/reference/form/mapping/keyscanner.hpp:55-110,
// keyscanner.cpp:29-91): a deque of recent scans (<= max_num_recent_scans), a
// deque of key scans; the oldest recent scan is promoted when its connection
// ratio exceeds keyscan_match_ratio, otherwise marginalised; key scans that stay
// unconnected for more than max_steps_unused_keyscan steps are marginalised, and
// the oldest key scan goes when there are more than max_num_keyscans.
#pragma once

#include "form/types.hpp"

#include <algorithm>
#include <cstdint>
#include <deque>
#include <functional>
#include <vector>

namespace form {

struct Scan {
  size_t idx;
  size_t unused_count = 0;
  size_t size = 0;
  Scan(size_t idx_, size_t size_ = 0) : idx(idx_), size(size_) {}
};

class KeyScanner {
public:
  struct Params {
    int64_t max_num_keyscans = 50;
    int64_t max_steps_unused_keyscan = 10;
    size_t max_num_recent_scans = 10;
    double keyscan_match_ratio = 0.1;
  };

  KeyScanner() : m_params() {}
  explicit KeyScanner(const Params &params) : m_params(params) {}

  /// Returns the scans to marginalise (keyscanner.cpp:29-91).
  std::vector<ScanIndex> step(ScanIndex idx, size_t size,
                              const std::function<size_t(ScanIndex)> &connections) noexcept {
    if (idx == 0) m_keyscans.emplace_back(idx, size);
    else m_recent_scans.emplace_back(idx, size);

    std::vector<ScanIndex> marg;
    // recent scans: promote or drop the one falling out of the window
    if (m_recent_scans.size() > m_params.max_num_recent_scans) {
      const Scan rf = m_recent_scans.front();
      m_recent_scans.pop_front();
      const double ratio = static_cast<double>(connections(rf.idx)) /
                           static_cast<double>(rf.size * m_recent_scans.size());
      if (ratio > m_params.keyscan_match_ratio) m_keyscans.push_back(rf);
      else marg.push_back(rf.idx);
    }
    // key scans: age out the ones no recent scan connects to
    std::vector<ScanIndex> finished;
    for (auto &kf : m_keyscans) {
      if (connections(kf.idx) > 0) kf.unused_count = 0;
      else ++kf.unused_count;
      // size_t > int64_t compares as unsigned in the reference (keyscanner.cpp:66): a negative
      // max_steps_unused_keyscan means "never"
      if (kf.unused_count > static_cast<size_t>(m_params.max_steps_unused_keyscan)) {
        marg.push_back(kf.idx);
        finished.push_back(kf.idx);
      }
    }
    m_keyscans.erase(std::remove_if(m_keyscans.begin(), m_keyscans.end(),
                                    [&](const Scan &f) {
                                      return std::find(finished.begin(), finished.end(), f.idx) !=
                                             finished.end();
                                    }),
                     m_keyscans.end());
    // hard cap
    if (m_params.max_num_keyscans > 0 &&
        static_cast<int64_t>(m_keyscans.size()) > m_params.max_num_keyscans) {
      marg.push_back(m_keyscans.front().idx);
      m_keyscans.pop_front();
    }
    return marg;
  }

  size_t size() const noexcept { return m_keyscans.size() + m_recent_scans.size(); }
  size_t newest_rf() const noexcept { return m_recent_scans.back().idx; }
  /// Oldest recent scan; the reference reads front() of a possibly empty deque
  /// on scan 0 (form.cpp:104-106) - here that case returns 0.
  size_t oldest_rf() const noexcept { return m_recent_scans.empty() ? 0 : m_recent_scans.front().idx; }
  size_t newest_kf() const noexcept { return m_keyscans.back().idx; }
  const std::deque<Scan> &keyscans() const noexcept { return m_keyscans; }
  const std::deque<Scan> &recent_scans() const noexcept { return m_recent_scans; }

private:
  Params m_params;
  std::deque<Scan> m_recent_scans;
  std::deque<Scan> m_keyscans;
};

} // namespace form
