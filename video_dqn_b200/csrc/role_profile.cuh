// Role profiling of the warp-specialised kernels (instrumented builds only).
#pragma once

namespace vdqn {

// Instrumented build (-DVDQN_ROLE_PROFILE, tools/role_profile.py): every role records the cycles it
// spent blocked on its two kinds of waits and its total loop time, per CTA -- the role that never
// waits is the bottleneck.  Compiles to nothing otherwise.
#ifdef VDQN_ROLE_PROFILE
static __device__ unsigned long long g_role_prof[160 * 16 * 4];   // one copy per translation unit
#define PROF_BEGIN long long prof_wa = 0, prof_wb = 0, prof_n = 0; const long long prof_t0 = clock64();
#define PROF_WAIT_A(x) { const long long w0_ = clock64(); x; prof_wa += clock64() - w0_; }
#define PROF_WAIT_B(x) { const long long w0_ = clock64(); x; prof_wb += clock64() - w0_; }
#define PROF_TILE ++prof_n;
#define PROF_END(role)                                                                   \
  if (lane == 0) {                                                                       \
    unsigned long long* p_ = g_role_prof + ((size_t)blockIdx.x * 16 + (role)) * 4;       \
    p_[0] = prof_wa; p_[1] = prof_wb; p_[2] = clock64() - prof_t0; p_[3] = prof_n;       \
  }
#else
#define PROF_BEGIN
#define PROF_WAIT_A(x) x;
#define PROF_WAIT_B(x) x;
#define PROF_TILE
#define PROF_END(role)
#endif


}  // namespace vdqn

// instrumented builds only (not part of include/vdqn.h): copies this translation unit's counters
#ifdef VDQN_ROLE_PROFILE
#define VDQN_DEFINE_ROLE_PROFILE_READER(name)                                                        \
  extern "C" int name(unsigned long long* out, int count) {                                          \
    cudaError_t e = cudaMemcpyFromSymbol(out, vdqn::g_role_prof, sizeof(unsigned long long) * count); \
    return e == cudaSuccess ? 0 : -1;                                                                 \
  }
#else
#define VDQN_DEFINE_ROLE_PROFILE_READER(name)                                        \
  extern "C" int name(unsigned long long* out, int count) { (void)out; (void)count; return -2; }
#endif
