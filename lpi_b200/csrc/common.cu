// Error reporting and device checks for liblpi_b200.so.
#include "lpi_internal.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace lpi {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(LPI_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return LPI_OK;
}

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("LPI_PDL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

}  // namespace lpi

extern "C" const char* lpi_last_error(void) { return lpi::g_err; }

extern "C" int lpi_version(void) { return 100; }

extern "C" int lpi_device_check(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return lpi::set_error(LPI_ERR_CUDA, "no CUDA device: %s", cudaGetErrorString(e));
    int major = 0, minor = 0, sms = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    if (major != 10) return lpi::set_error(LPI_ERR_UNSUPPORTED, "lpi_b200 needs an sm_100a device, found sm_%d%d", major, minor);
    return LPI_OK;
}
