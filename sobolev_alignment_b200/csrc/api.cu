// Error state and version of libsobolev_b200 (the C ABI declared in include/sobolev_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/sobolev_b200.h"
#include "common.cuh"

namespace sab {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

}  // namespace sab

// ---------------------------------------------------------------------------------------------
// tuning / diagnostic options: an explicit sa_set_option() value wins, otherwise the SAB_* environment
// variable of the same name, read ONCE (at its first use) and cached.
// ---------------------------------------------------------------------------------------------
namespace sab {
namespace {
struct Opt {
  char name[40];
  char value[64];
  bool has_value;
};
constexpr int MAX_OPTS = 48;
Opt g_opts[MAX_OPTS];
int g_nopts = 0;
std::mutex g_opt_mu;

Opt* find_or_add(const char* name, bool from_env) {
  for (int i = 0; i < g_nopts; ++i)
    if (std::strcmp(g_opts[i].name, name) == 0) return &g_opts[i];
  if (g_nopts == MAX_OPTS) return nullptr;
  Opt* o = &g_opts[g_nopts++];
  std::snprintf(o->name, sizeof(o->name), "%s", name);
  o->has_value = false;
  if (from_env) {
    const char* v = std::getenv(name);
    if (v) {
      std::snprintf(o->value, sizeof(o->value), "%s", v);
      o->has_value = true;
    }
  }
  return o;
}
}  // namespace

bool option(const char* name, char (&out)[64]) {
  std::lock_guard<std::mutex> lk(g_opt_mu);
  Opt* o = find_or_add(name, true);
  if (o == nullptr || !o->has_value) return false;
  std::memcpy(out, o->value, sizeof(o->value));
  return true;
}
int option_int(const char* name, int dflt) {
  char buf[64];
  return option(name, buf) ? std::atoi(buf) : dflt;
}
}  // namespace sab

extern "C" int sa_set_option(const char* name, const char* value) {
  if (name == nullptr || std::strncmp(name, "SAB_", 4) != 0) {
    sab::set_error("option names start with SAB_");
    return 2;
  }
  std::lock_guard<std::mutex> lk(sab::g_opt_mu);
  sab::Opt* o = sab::find_or_add(name, false);
  if (o == nullptr) {
    sab::set_error("too many options");
    return 2;
  }
  if (value == nullptr) {   // back to the environment's value
    const char* v = std::getenv(name);
    o->has_value = v != nullptr;
    if (v) std::snprintf(o->value, sizeof(o->value), "%s", v);
  } else {
    std::snprintf(o->value, sizeof(o->value), "%s", value);
    o->has_value = true;
  }
  return 0;
}

extern "C" const char* sa_last_error(void) { return sab::g_error; }
extern "C" int sa_version(void) { return 100; }
