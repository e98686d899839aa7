/*
 * oracle/ref_shim/ref_prelude.h -- force-included (-include) in front of every translation unit of
 * oracle/_ref, the reference's own sources among them.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference seeds three private std::mt19937 from std::random_device (sampler.cpp:60,
 * chainarray.cpp:32-34,57-59), so its runs cannot be reproduced or compared as they stand.  Without touching
 * the sources, this header (1) pins the seed: `random_device` is redirected to a functor that returns the
 * harness's seed, and (2) records every variate the reference draws: `normal_distribution` and
 * `uniform_real_distribution` are redirected to wrappers that hold the real libstdc++ distribution, draw from
 * the reference's real mt19937 and report the value to the harness, which attributes it to (round, chain)
 * and hands the tables to the oracle's replay mode.  <random> is included first, so libstdc++ itself is
 * compiled with the real names.
 */
#ifndef SLREF_PRELUDE_H
#define SLREF_PRELUDE_H
#ifdef __cplusplus
#include <random>

namespace slref {
extern unsigned g_seed;
void log_draw(int kind /*0 normal, 1 uniform*/, const void* dist, double v);

struct fixed_random_device {
  typedef unsigned result_type;
  unsigned operator()() { return g_seed; }
};
template <class T = double>
struct logged_normal {
  std::normal_distribution<T> d;
  template <class G>
  T operator()(G& g) {
    const T v = d(g);
    log_draw(0, this, (double)v);
    return v;
  }
};
template <class T = double>
struct logged_uniform {
  std::uniform_real_distribution<T> d;
  template <class G>
  T operator()(G& g) {
    const T v = d(g);
    log_draw(1, this, (double)v);
    return v;
  }
};
}  // namespace slref

namespace std {
using slref::fixed_random_device;
using slref::logged_normal;
using slref::logged_uniform;
}  // namespace std

/* comms/socket.cpp draws its ZeroMQ identity from random_device (socket.cpp:188-198): pinned, every worker socket of a
 * process would get the same identity and the delegator's ROUTER would talk to the first one only.  That translation
 * unit is compiled with -DSLREF_REAL_RANDOM_DEVICE (oracle/Makefile.full). */
#ifndef SLREF_REAL_RANDOM_DEVICE
#define random_device fixed_random_device
#endif
#define normal_distribution logged_normal
#define uniform_real_distribution logged_uniform
#endif
#endif
