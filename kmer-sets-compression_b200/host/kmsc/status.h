// kmsc/status.h -- Status / StatusOr with the slice of absl::Status the reference's
// API surface uses (ok(), status(), value(), message(), ToString()); error texts of
// the reference are kept (lib/core/io.h:27,63-65; lib/core/kmer_counter.h:163-203).
#ifndef KMSC_HOST_STATUS_H_
#define KMSC_HOST_STATUS_H_
#include <optional>
#include <string>
#include <utility>

namespace kmsc {

enum class StatusCode { kOk = 0, kFailedPrecondition = 9, kInternal = 13 };

class Status {
 public:
  Status() = default;
  Status(StatusCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  bool ok() const { return code_ == StatusCode::kOk; }
  StatusCode code() const { return code_; }
  const std::string& message() const { return msg_; }
  std::string ToString() const {
    if (ok()) return "OK";
    return std::string(code_ == StatusCode::kInternal ? "INTERNAL: " : "FAILED_PRECONDITION: ") + msg_;
  }

 private:
  StatusCode code_ = StatusCode::kOk;
  std::string msg_;
};

inline Status OkStatus() { return Status(); }
inline Status InternalError(std::string m) { return Status(StatusCode::kInternal, std::move(m)); }
inline Status FailedPreconditionError(std::string m) { return Status(StatusCode::kFailedPrecondition, std::move(m)); }

template <typename T>
class StatusOr {
 public:
  StatusOr(const Status& s) : status_(s) {}
  StatusOr(Status&& s) : status_(std::move(s)) {}
  StatusOr(const T& v) : value_(v) {}
  StatusOr(T&& v) : value_(std::move(v)) {}
  bool ok() const { return status_.ok(); }
  const Status& status() const { return status_; }
  T& value() & { return *value_; }
  const T& value() const& { return *value_; }
  T&& value() && { return std::move(*value_); }
  T& operator*() { return *value_; }
  T* operator->() { return &*value_; }

 private:
  Status status_;
  std::optional<T> value_;
};

}  // namespace kmsc
#endif
