// lcpc_b200/csrc/tunables.cpp -- named integer knobs for A/B measurements.
// A knob's value is, in this order: what lcpc_b200_set_tunable() stored, the environment variable
// LCPC_B200_<NAME>, the default the call site passes.  Knobs never change results, only schedules.
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>

#include "kernels.h"

namespace lcpc {

namespace {
std::mutex g_mu;
std::map<std::string, long> g_set;
}  // namespace

long tunable(const char *name, long dflt) {
  {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_set.find(name);
    if (it != g_set.end()) return it->second;
  }
  std::string env = std::string("LCPC_B200_") + name;
  const char *e = getenv(env.c_str());
  return (e && *e) ? atol(e) : dflt;
}

void set_tunable(const char *name, long value) {
  std::lock_guard<std::mutex> g(g_mu);
  g_set[name] = value;
}

void clear_tunable(const char *name) {
  std::lock_guard<std::mutex> g(g_mu);
  g_set.erase(name);
}

}  // namespace lcpc
