/*
 * glog/logging.h shim — TEST INFRASTRUCTURE ONLY (oracle/_ref).  The reference sources use
 * LOG(INFO) (camera_model.cpp:51, lidar_model.cpp:49) and LOG(FATAL)
 * (module_factory_registry.h:47,66).  INFO/WARNING/ERROR go to stderr; FATAL throws
 * std::runtime_error instead of aborting so that a test process survives to report it.
 */
#ifndef A3D_ORACLE_GLOG_SHIM_H_
#define A3D_ORACLE_GLOG_SHIM_H_

#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>

namespace a3d_glog_shim {
enum Severity { INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3 };

class LogMessage {
 public:
  LogMessage(Severity s, const char* file, int line) : s_(s) { ss_ << file << ":" << line << "] "; }
  ~LogMessage() noexcept(false) {
    if (s_ == FATAL) throw std::runtime_error(ss_.str());
    if (s_ != INFO || verbose()) std::cerr << ss_.str() << std::endl;
  }
  std::ostream& stream() { return ss_; }
  static bool& verbose() {
    static bool v = false;
    return v;
  }

 private:
  Severity s_;
  std::ostringstream ss_;
};
}  // namespace a3d_glog_shim

#define LOG(severity) ::a3d_glog_shim::LogMessage(::a3d_glog_shim::severity, __FILE__, __LINE__).stream()
#define CHECK(cond) \
  if (!(cond)) LOG(FATAL) << "Check failed: " #cond " "
#define CHECK_NOTNULL(p) (p)

#endif  // A3D_ORACLE_GLOG_SHIM_H_
