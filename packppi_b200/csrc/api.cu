// Library-level entry points: version, error text, packed-weight layout, device check.
#include "common.cuh"
#include "weights_layout.h"

namespace pp {
thread_local char g_last_error[512] = "";
}

extern "C" int pp_abi_version() { return 2; }

extern "C" const char* pp_last_error() { return pp::g_last_error; }

extern "C" int64_t pp_layout_count() { return pp::wl::NUM_ENTRIES; }

extern "C" int64_t pp_layout_total_floats() { return pp::wl::kTotalFloats; }

extern "C" int pp_layout_entry(int64_t i, const char** name, int64_t* offset, int64_t* size) {
  PP_REQUIRE(i >= 0 && i < pp::wl::NUM_ENTRIES, "entry out of range");
  *name = pp::wl::kInfo[i].name;
  *offset = pp::wl::offset_of((int)i);
  *size = pp::wl::kInfo[i].size;
  return 0;
}

extern "C" int64_t pp_geo_stride() { return PP_GEO_STRIDE; }
extern "C" int64_t pp_table_stride() { return PP_TBL_STRIDE; }

// 0 when the current device can run this library (compute capability 10.x), else an error is recorded.
extern "C" int pp_check_device() {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    snprintf(pp::g_last_error, sizeof(pp::g_last_error), "pp_check_device: no CUDA device");
    return 1;
  }
  if (prop.major != 10) {
    snprintf(pp::g_last_error, sizeof(pp::g_last_error),
             "pp_check_device: %s is sm_%d%d; libpackppi_b200 is built for sm_100a only", prop.name, prop.major, prop.minor);
    return 1;
  }
  return 0;
}
