// host_tables.hpp — host-side tables shared by the engine and the host-compiled test harness.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../include/alpharat_cuda.h"

namespace ar_host {

// calculate_collisions_left (search.rs:437-450): libm powf and round-half-away, evaluated on the
// host once per configuration and indexed by node_count on the device.
inline uint32_t collisions_left(uint32_t n, const ar_search_cfg& c) {
  if (n >= c.collision_scaling_end) return c.collision_limit_max;
  if (n <= c.collision_scaling_start) return c.collision_limit_min;
  float ratio = (float)(n - c.collision_scaling_start) /
                (float)(c.collision_scaling_end - c.collision_scaling_start);
  float scaled = (float)c.collision_limit_min +
                 ((float)c.collision_limit_max - (float)c.collision_limit_min) *
                     powf(ratio, c.collision_scaling_power);
  float r = roundf(scaled);
  uint32_t v = !(r > 0.0f) ? 0u : (r >= 4294967296.0f ? 0xffffffffu : (uint32_t)r);
  return std::min(std::max(v, c.collision_limit_min), c.collision_limit_max);
}

inline std::vector<uint16_t> collision_table(const ar_search_cfg& c, uint32_t len) {
  std::vector<uint16_t> t(len);
  for (uint32_t n = 0; n < len; ++n) t[n] = (uint16_t)std::min<uint32_t>(collisions_left(n, c), 65535u);
  return t;
}

}  // namespace ar_host
