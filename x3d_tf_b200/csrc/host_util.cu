// Host-side pieces of the C ABI: error text, version, CRC-32C.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tma_common.cuh"

namespace x3d {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return X3D_ERR_LAUNCH;
  }
  return X3D_OK;
}

// CRC-32C, slicing-by-8 (tables built on first use; thread-safe via static init).
struct CrcTables {
  uint32_t t[8][256];
  CrcTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
};

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("X3D_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// Properties of the CURRENT device, cached per device ordinal (one process may drive several GPUs).
struct DevInfo { int sms = -1, smem = 0, major = 0; };
static DevInfo g_dev[kMaxDevices];
static const DevInfo& query_device() {
  static const DevInfo none{0, 0, 0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return none;
  DevInfo& d = g_dev[dev];
  if (d.sms < 0) {
    int sms = 0;
    cudaDeviceGetAttribute(&d.smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    d.sms = sms;
  }
  return d;
}
int device_sm_count() { return query_device().sms; }
int device_max_smem() { return query_device().smem; }
bool device_is_sm100() { return query_device().major == 10; }

}  // namespace x3d

extern "C" {

int x3d_version(void) { return X3D_B200_VERSION; }

const char* x3d_last_error(void) { return x3d::g_err; }

uint32_t x3d_crc32c(const void* data, size_t len, uint32_t crc) {
  static const x3d::CrcTables T;
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = ~crc;
  while (len && (reinterpret_cast<uintptr_t>(p) & 7)) {
    c = T.t[0][(c ^ *p++) & 0xff] ^ (c >> 8);
    --len;
  }
  while (len >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = T.t[7][w & 0xff] ^ T.t[6][(w >> 8) & 0xff] ^ T.t[5][(w >> 16) & 0xff] ^
        T.t[4][(w >> 24) & 0xff] ^ T.t[3][(w >> 32) & 0xff] ^ T.t[2][(w >> 40) & 0xff] ^
        T.t[1][(w >> 48) & 0xff] ^ T.t[0][(w >> 56) & 0xff];
    p += 8;
    len -= 8;
  }
  while (len--) c = T.t[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return ~c;
}

}  // extern "C"
