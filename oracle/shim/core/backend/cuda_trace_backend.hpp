// oracle/shim/core/backend/cuda_trace_backend.hpp — TEST INFRASTRUCTURE (oracle/Makefile target `refb200`).
//
// Placed on the include path BEFORE the reference tree when the reference's own, unmodified simulator.cpp is compiled
// with -DLUMICE_CUDA_ENABLED: its `#include "core/backend/cuda_trace_backend.hpp"` then resolves to this file, and the
// name the reference driver instantiates -- CreateBackend(): std::make_unique<CudaTraceBackend>(&logger),
// simulator.cpp:854-919 -- is this repo's adapter (adapter/b200_trace_backend.hpp over the C ABI). Nothing of the
// reference is patched or copied: Simulator::Run, SimulateOneWavelengthWithBackend and the third-clock drain
// (DrainDeviceXyz, simulator.cpp:1409-1477) run as shipped and drive the B200 engine through the TraceBackend seam.
// A maintainer's real integration registers a backend kind of its own instead.
#ifndef ORACLE_SHIM_CUDA_TRACE_BACKEND_HPP_
#define ORACLE_SHIM_CUDA_TRACE_BACKEND_HPP_

#include <cstdlib>
#include <string>
#include <vector>

#include "../../../../adapter/b200_trace_backend.hpp"

namespace lumice {

class Logger;

// Devices the backend fans a session out to: HALOTRACE_B200_DEVICES="0,1,2,3" (default: device 0).
inline std::vector<int> ShimDevices() {
  std::vector<int> out;
  const char* env = std::getenv("HALOTRACE_B200_DEVICES");
  std::string s = env ? env : "0";
  size_t pos = 0;
  while (pos < s.size()) {
    size_t comma = s.find(',', pos);
    if (comma == std::string::npos) comma = s.size();
    if (comma > pos) out.push_back(std::atoi(s.substr(pos, comma - pos).c_str()));
    pos = comma + 1;
  }
  if (out.empty()) out.push_back(0);
  return out;
}

inline bool CudaDeviceAvailable() {
  HbEngine* probe = nullptr;
  if (hb_create(ShimDevices()[0], &probe) != HB_OK) return false;
  hb_destroy(probe);
  return true;
}
inline std::string CudaDeviceDiagnostics() { return "halotrace-b200 shim (oracle/shim): B200TraceBackend behind the CudaTraceBackend name"; }

class CudaTraceBackend : public B200TraceBackend {
 public:
  explicit CudaTraceBackend(Logger* = nullptr) : B200TraceBackend(ShimDevices()) {}
};

}  // namespace lumice

#endif  // ORACLE_SHIM_CUDA_TRACE_BACKEND_HPP_
