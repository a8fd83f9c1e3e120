// TEST INFRASTRUCTURE ONLY (oracle/). std-only stand-in for absl::StatusOr.
#ifndef KMSC_ORACLE_SHIM_STATUSOR_H_
#define KMSC_ORACLE_SHIM_STATUSOR_H_
#include <optional>
#include <utility>
#include "absl/status/status.h"
namespace absl {
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
}  // namespace absl
#endif
