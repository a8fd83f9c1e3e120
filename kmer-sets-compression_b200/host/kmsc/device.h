// kmsc/device.h -- RAII over the C ABI (include/kmsc.h): one process-wide context
// per GPU and shared, immutable device-set handles. Errors from the library are
// fatal for the value-returning compute paths, exactly like the reference's pure
// compute paths "cannot fail" (SURVEY 8b); IO / parse paths return Status.
#ifndef KMSC_HOST_DEVICE_H_
#define KMSC_HOST_DEVICE_H_
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <string>

#include "kmsc.h"

namespace kmsc {

class Device {
 public:
  // The context of the GPU selected by KMSC_DEVICE (default 0). There is no CPU
  // fallback: if the library cannot get a device this aborts with its message.
  static kmsc_ctx* Ctx() {
    static Device d;
    return d.ctx_;
  }
  // serialises callers: a kmsc_ctx is driven by one host thread at a time
  static std::mutex& Mu() {
    static std::mutex m;
    return m;
  }
  static void Check(int rc, const char* what) {
    if (rc != KMSC_OK) {
      std::fprintf(stderr, "libkmsc: %s failed (%d): %s\n", what, rc, kmsc_last_error());
      std::abort();
    }
  }

 private:
  Device() {
    const char* e = std::getenv("KMSC_DEVICE");
    Check(kmsc_ctx_create(e ? std::atoi(e) : 0, nullptr, &ctx_), "kmsc_ctx_create");
  }
  ~Device() { kmsc_ctx_destroy(ctx_); }
  kmsc_ctx* ctx_ = nullptr;
};

// shared ownership of an immutable device set
struct SetHandle {
  explicit SetHandle(kmsc_set* s) : set(s) {}
  ~SetHandle() {
    std::lock_guard<std::mutex> l(Device::Mu());
    kmsc_set_free(Device::Ctx(), set);
  }
  SetHandle(const SetHandle&) = delete;
  SetHandle& operator=(const SetHandle&) = delete;
  kmsc_set* set;
};
using SetPtr = std::shared_ptr<SetHandle>;
inline SetPtr MakeSetPtr(kmsc_set* s) { return std::make_shared<SetHandle>(s); }

}  // namespace kmsc
#endif
