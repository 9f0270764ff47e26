// b200_implementation.h — `simdutf::b200::implementation`, the new entry of simdutf's implementation list.
//
// Modelled on the reference's per-kernel headers (e.g. src/simdutf/haswell/implementation.h:13-19): a final
// subclass of simdutf::implementation (reference include/simdutf/implementation.h:3302-5066) named "b200".
// Every pure virtual of this reference forwards to one b200_host_* entry point of the C ABI
// (include/simdutf_b200.h); a virtual a later reference adds would get the reference's "unsupported" value
// (generated, tools/gen_b200_cxx.py; empty today).  C++11-clean: it is #included by the reference's unity translation unit.
#ifndef SIMDUTF_B200_IMPLEMENTATION_H
#define SIMDUTF_B200_IMPLEMENTATION_H

#include "simdutf/implementation.h"

namespace simdutf {
namespace b200 {

using namespace simdutf;

class implementation final : public simdutf::implementation {
public:
  // required_instruction_sets is a constructor argument in the reference; ours depends on the machine, so the
  // (virtual) accessor is overridden instead — see below.
  simdutf_really_inline implementation() : simdutf::implementation("b200", "NVIDIA B200 (sm_100a) CUDA kernels", 0) {}

  // supported_by_runtime_system() is non-virtual: (detected & required) == required
  // (reference src/implementation.cpp:35-41).  With at least one sm_100 device we require nothing; without one
  // we require a bit no CPU detector ever reports, so tests/benchmarks skip "b200" on GPU-less hosts
  // (reference tests/helpers/test.cpp:165-169) instead of failing.
  uint32_t required_instruction_sets() const override;

#include "b200_decls.inc"

#ifdef SIMDUTF_INTERNAL_TESTS
  // Developer hook of the reference (include/simdutf/implementation.h:5016-5037, tests/internal_tests.cpp): what the
  // public API cannot reach — the host path spread over every device, the per-thread host paths, the shard helpers.
  std::vector<TestProcedure> internal_tests() const override;
#endif
};

} // namespace b200
} // namespace simdutf

#endif // SIMDUTF_B200_IMPLEMENTATION_H
