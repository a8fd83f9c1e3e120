// stubs.cu -- entry points of include/kmsc.h whose kernels are not built yet.
// They fail loudly (KMSC_E_STATE); nothing here computes on the CPU.
#include "kmsc_common.cuh"
using namespace kmsc;
#define KMSC_STUB(name) set_error(name ": not built yet in this revision"); return KMSC_E_STATE;
extern "C" {
int kmsc_codec_encode(kmsc_ctx*, const kmsc_set*, uint8_t**, int64_t*) { KMSC_STUB("kmsc_codec_encode") }
int kmsc_codec_decode(kmsc_ctx*, const uint8_t*, int64_t, kmsc_set**) { KMSC_STUB("kmsc_codec_decode") }
}
