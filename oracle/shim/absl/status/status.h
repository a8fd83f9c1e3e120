// TEST INFRASTRUCTURE ONLY (oracle/). std-only stand-in for absl::Status.
#ifndef KMSC_ORACLE_SHIM_STATUS_H_
#define KMSC_ORACLE_SHIM_STATUS_H_
#include <string>
#include <utility>
namespace absl {
enum class StatusCode { kOk = 0, kInternal = 13, kFailedPrecondition = 9 };
class Status {
 public:
  Status() = default;
  Status(StatusCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  bool ok() const { return code_ == StatusCode::kOk; }
  StatusCode code() const { return code_; }
  const std::string& message() const { return msg_; }
  std::string ToString() const {
    if (ok()) return "OK";
    return (code_ == StatusCode::kInternal ? "INTERNAL: " : "FAILED_PRECONDITION: ") + msg_;
  }
 private:
  StatusCode code_ = StatusCode::kOk;
  std::string msg_;
};
inline Status OkStatus() { return Status(); }
inline Status InternalError(std::string m) { return Status(StatusCode::kInternal, std::move(m)); }
inline Status FailedPreconditionError(std::string m) {
  return Status(StatusCode::kFailedPrecondition, std::move(m));
}
}  // namespace absl
#endif
